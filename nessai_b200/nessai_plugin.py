"""Bindings into the real ``nessai`` package (plugin points P2 / P3).

Importing this module requires ``nessai`` to be importable.  It defines

* ``B200NessaiFlowProposal(nessai.proposal.FlowProposal)`` with
  ``_FlowModelClass = B200FlowModel`` -- every flow evaluation of the unmodified
  reference proposal (train / forward_pass / backward_pass / truncation rules)
  runs on the CUDA kernels -- and a ``populate`` override that runs the fused
  device loop whenever the configuration allows it (``parameter_maps``: every
  reparameterisation name the reference registers -- z-score / null / scale /
  rescale-to-bounds incl. the named pre- / post-rescalings and boundary inversion,
  ``Angle``, ``ToCartesian``, ``AnglePair``, ``Dequantise``; ``map_to_unit_hypercube``; the
  three truncation rules; weight accumulation) and otherwise defers to the reference's
  host loop;
* ``B200AugmentedFlowProposal``: the same for ``AugmentedFlowProposal``;
* the entry point ``nessai.proposals: b200flowproposal``, so
  ``FlowSampler(model, flow_proposal_class="b200flowproposal")`` picks it up
  through ``nessai.proposal.utils.get_flow_proposal_class``
  (/root/reference/src/nessai/proposal/utils.py:110-155).
"""

from __future__ import annotations

import datetime
import logging
from typing import NamedTuple

import numpy as np
from nessai import config as nessai_config
from nessai.livepoint import empty_structured_array as nessai_empty_structured_array
from nessai.model import Model as _ReferenceModel
from nessai.proposal.augmented import AugmentedFlowProposal
from nessai.proposal.flowproposal import FlowProposal
from nessai.reparameterisations import (
    Angle,
    AnglePair,
    Dequantise,
    NullReparameterisation,
    RescaleToBounds,
    ScaleAndShift,
    ToCartesian,
)
from nessai.utils import rescaling as _ref_rescaling

from .flowmodel import B200FlowModel
from .proposal import GeneralPopulateEngine, IndexPool, PopulateEngine, detect_uniform_box_prior

logger = logging.getLogger(__name__)


# per-parameter map kinds of the device tail (csrc/reparam_tail.cuh: TailKind)
(KIND_IDENTITY, KIND_SIGMOID, KIND_ABS, KIND_EXP, KIND_LOG, KIND_NORMAL_CDF,
 KIND_NORMAL_QUANTILE, KIND_ANGLE, KIND_ANGLE_MOD, KIND_RADIUS, KIND_RADIUS_CHI, KIND_FLOOR,
 KIND_ANGLE_ABS, KIND_ZENITH, KIND_DECLINATION, KIND_RADIUS3, KIND_RADIUS3_CHI, KIND_GAUSS_AUX) = range(18)
KIND_FLOOR_AFTER = 0x100  # flag: floor of the final value (Dequantise with a post-rescaling)

# the INVERSE function of a named rescaling (utils/rescaling.py:410-417) -> the kind of h
_INVERSE_KINDS = (
    (_ref_rescaling.sigmoid, KIND_SIGMOID),
    (_ref_rescaling.exp_with_log_jacobian, KIND_EXP),
    (_ref_rescaling.log_with_log_jacobian, KIND_LOG),
    (_ref_rescaling.gaussian_cdf, KIND_NORMAL_CDF),
    (_ref_rescaling.inverse_gaussian_cdf, KIND_NORMAL_QUANTILE),
)


def _kind_of(fn):
    for ref_fn, kind in _INVERSE_KINDS:
        if fn is ref_fn:
            return kind
    return None


class ParameterMaps(NamedTuple):
    """``x = h(pre_scale * x'[src] + pre_shift) * scale + shift`` per x-space parameter ``names[d]``
    (``src[d]``: the one, two -- ``Angle``, ``ToCartesian`` -- or three -- ``AnglePair`` -- prime
    parameters it reads)."""

    kind: np.ndarray
    scale: np.ndarray
    shift: np.ndarray
    pre_scale: np.ndarray
    pre_shift: np.ndarray
    src: np.ndarray  # (D, 3) indices into the prime parameters
    names: list  # x-space parameters: the model's, then auxiliary ones (flowproposal/base.py:539-546)

    @property
    def affine(self) -> bool:
        d = np.arange(len(self.kind))
        return not np.any(self.kind != KIND_IDENTITY) and bool(np.all(self.src[:, 0] == d))

    @property
    def has_pre_affine(self) -> bool:
        return bool(np.any(self.pre_scale != 1.0) or np.any(self.pre_shift != 0.0))


def parameter_maps(rep, prime_parameters, model_names, x_parameters=None):
    """The combined reparameterisation ``rep`` (reparameterisations/combined.py:154-192) as
    per-parameter maps ``x = h(a x' + b) * scale + shift`` with ``h`` = identity / sigmoid /
    ``|.|`` / exp / log / normal CDF / normal quantile, plus the Cartesian pairs of ``Angle`` (a
    ``ParameterMaps``), or ``None`` when it is not made of such maps.  ``x_parameters``: the
    proposal's x-space parameters (model names, then auxiliary ones; default the model's).
    ``log|J| = sum log|scale| + sum log|a| + sum log|h'|``.  Recognised:

    * ``NullReparameterisation`` (reparameterisations/null.py): identity;
    * ``ScaleAndShift`` / ``Rescale`` (rescale.py:233-291): ``x = x' * scale + shift``; with a
      named post-rescaling ``x = Q^-1(x') * scale + shift`` ("zscore-gaussian-cdf"); with a
      named pre-rescaling ``x = P^-1(x' * scale + shift)`` ("z-score-logit", "log-z-score",
      "z-score-inv-gaussian-cdf": ``a = scale``, ``b = shift``);
    * the stock ``RescaleToBounds`` (rescale.py:321-731):
      ``x = (hi - lo) * (h(x') - r0) / (r1 - r0) + lo + offset`` with ``[lo, hi]`` the current
      (possibly data-updated) bounds and ``[r0, r1]`` the rescale bounds (rescale.py:533-553,
      669-680), i.e. ``scale = (hi - lo) / (r1 - r0)``, ``shift = lo + offset - scale * r0``;
      ``h`` from a named post-rescaling ("logit" -> sigmoid, "log" -> exp, ...;
      utils/rescaling.py:290-417); with boundary inversion (rescale.py:570-590) ``h = |.|`` and
      the detected edge picks the map: "lower" ``[0, 1] -> [lo, hi]``, "upper" ``1 - |x'|`` first
      (negative scale), no edge ``[-1, 1] -> [lo, hi]`` with ``h`` = identity; with a named
      pre-rescaling that affine map becomes ``a, b`` and ``h = P^-1`` comes last;
    * the stock ``Angle`` (angle.py:17-186; "angle", "angle-pi", "angle-2pi", "periodic"): the two
      prime parameters are Cartesian coordinates, ``angle = atan2(y', x') [% 2 pi] / scale`` and
      ``r = sqrt(x'^2 + y'^2)`` with ``log|J| -= log r``; ``r`` is the model's radial parameter or an
      auxiliary one with a ``chi(2)`` prior.

    * the stock ``ToCartesian`` (angle.py:189-232): a non-periodic parameter as the angle of a
      Cartesian pair, inverse ``|atan2(y', x') / scale| * (hi - lo) + lo``, ``log|J| += log(hi - lo)``;
    * the stock ``AnglePair`` (angle.py:235-538): two angles (+ a radial parameter, or an auxiliary
      one with a ``chi(3)`` prior) as a Cartesian triple, conventions "az-zen" and "ra-dec";
    * the stock ``Dequantise`` (discrete.py): ``RescaleToBounds`` over ``[lo, hi + 1]`` whose
      pre-rescaling inverse is ``floor``.

    Anything else (user rescaling callables, two non-linear stages on one parameter, an edge
    not detected yet, user classes) keeps the reference's host loop."""
    if rep is None:
        return None
    prime_parameters = list(prime_parameters)
    x_parameters = list(model_names) if x_parameters is None else list(x_parameters)
    if prime_parameters != [p for r in rep.values() for p in r.output_parameters]:
        return None
    specs = {}  # x-space name -> (kind, scale, shift, a, b, source prime names)
    for r in rep.values():
        if isinstance(r, NullReparameterisation):
            for p, pp in zip(r.parameters, r.output_parameters):
                specs[p] = (KIND_IDENTITY, 1.0, 0.0, 1.0, 0.0, (pp, pp, pp))
            continue
        if type(r) in (Angle, ToCartesian):
            if len(r.output_parameters) != 2 or not np.isfinite(r.scale) or r.scale == 0:
                return None
            pair = (r.x, r.y, r.x)
            if type(r) is ToCartesian:
                lo, hi = (float(b) for b in r.prior_bounds[r.angle])
                if not (np.isfinite(lo) and np.isfinite(hi) and hi > lo):
                    return None
                specs[r.angle] = (KIND_ANGLE_ABS, hi - lo, lo, 1.0 / float(r.scale), 0.0, pair)
            else:
                specs[r.angle] = (KIND_ANGLE_MOD if r._zero_bound else KIND_ANGLE, 1.0 / float(r.scale), 0.0, 1.0, 0.0,
                                  pair)
            specs[r.radial] = (KIND_RADIUS_CHI if r.chi else KIND_RADIUS, 1.0, 0.0, 1.0, 0.0, pair)
            continue
        if type(r) is AnglePair:
            if len(r.output_parameters) != 3 or r.convention not in ("az-zen", "ra-dec"):
                return None
            triple = (r.x, r.y, r.z)
            specs[r.angles[0]] = (KIND_ANGLE_MOD if r._modulo_2pi else KIND_ANGLE, 1.0, 0.0, 1.0, 0.0, triple)
            specs[r.angles[1]] = (KIND_ZENITH if r.convention == "az-zen" else KIND_DECLINATION, 1.0, 0.0, 1.0, 0.0,
                                  triple)
            specs[r.radial] = (KIND_RADIUS3_CHI if r.chi else KIND_RADIUS3, 1.0, 0.0, 1.0, 0.0, triple)
            continue
        if not (isinstance(r, ScaleAndShift) or type(r) in (RescaleToBounds, Dequantise)) or not r.one_to_one:
            # (RescaleToBounds subclasses may override the rescaling hooks: stock classes only)
            return None
        if type(r) is Dequantise:
            pre = KIND_FLOOR  # discrete.py:77-78
        else:
            pre = _kind_of(r.pre_rescaling_inv) if r.has_pre_rescaling else KIND_IDENTITY
        post = _kind_of(r.post_rescaling_inv) if r.has_post_rescaling else KIND_IDENTITY
        floor_after = 0
        if pre == KIND_FLOOR and post not in (None, KIND_IDENTITY):
            # "dequantise-logit": x = floor(Q^-1(x') * scale + shift) -- floor has no log-Jacobian, so it is a
            # flag on the post-rescaling's kind rather than a second stage
            pre, floor_after = KIND_IDENTITY, KIND_FLOOR_AFTER
        if pre is None or post is None or (pre != KIND_IDENTITY and post != KIND_IDENTITY):
            return None
        for p, pp in zip(r.parameters, r.output_parameters):
            k = post
            if isinstance(r, ScaleAndShift):
                s, t = float(r.scale[p]), (float(r.shift[p]) if r.shift else 0.0)
            else:
                lo, hi = (float(b) for b in r.bounds[p])
                off = float(r.offsets[p])
                if r.boundary_inversion and p in r.boundary_inversion:
                    edge = (r._edges or {}).get(p)
                    if post != KIND_IDENTITY or pre != KIND_IDENTITY or edge is None:
                        return None
                    if edge == "lower":
                        k, s, t = KIND_ABS, hi - lo, lo + off
                    elif edge == "upper":
                        k, s, t = KIND_ABS, -(hi - lo), hi + off
                    elif not edge:
                        s, t = (hi - lo) / 2.0, lo + off + (hi - lo) / 2.0
                    else:
                        return None
                else:
                    s = (hi - lo) / float(r._rescale_factor[p])
                    t = lo + off - s * float(r._rescale_shift[p])
            if pre != KIND_IDENTITY:
                specs[p] = (pre, 1.0, 0.0, s, t, (pp, pp, pp))  # the affine map first, then P^-1
            else:
                specs[p] = (k | floor_after, s, t, 1.0, 0.0, (pp, pp, pp))
    if set(specs) != set(x_parameters) or len(x_parameters) != len(prime_parameters):
        return None
    if list(x_parameters[: len(model_names)]) != list(model_names):
        return None
    try:
        src = np.asarray([[prime_parameters.index(q) for q in specs[n][5]] for n in x_parameters], dtype=np.int32)
    except ValueError:
        return None
    if set(src.ravel().tolist()) != set(range(len(prime_parameters))):
        return None  # a prime parameter nobody reads: not a bijection
    cols = list(zip(*(specs[n][:5] for n in x_parameters)))
    maps = ParameterMaps(np.asarray(cols[0], dtype=np.int32), *(np.asarray(c, dtype=np.float64) for c in cols[1:]),
                         src, x_parameters)
    for v in maps[1:5]:
        if not np.all(np.isfinite(v)):
            return None
    if np.any(maps.scale == 0.0) or np.any(maps.pre_scale == 0.0):
        return None
    return maps


def diagonal_rescaling(rep, prime_parameters, model_names):
    """``(scale, shift)`` when ``parameter_maps`` finds a diagonal affine (every ``h`` the
    identity): what the fused populate tail of the draw kernels evaluates itself."""
    maps = parameter_maps(rep, prime_parameters, model_names)
    if maps is None or not maps.affine:
        return None
    return maps.scale, maps.shift


class B200NessaiFlowProposal(FlowProposal):
    """``FlowProposal`` whose flow and populate loop run on a B200.

    Extra keyword argument ``device_prior``: ``"auto"`` (default) fuses the prior
    into the device loop when the model's prior is a uniform box (checked
    numerically, or declared with ``model.uniform_box_prior``); ``False`` always
    evaluates ``model.batch_evaluate_log_prior`` on the host.
    """

    _FlowModelClass = B200FlowModel

    def __init__(self, model, device_prior="auto", **kwargs):
        super().__init__(model, **kwargs)
        self.device_prior = device_prior
        self._engine = None
        self._log_prior_const = None

    # ------------------------------------------------------------ eligibility
    def _parameter_maps(self):
        """(kind, scale, shift) in prime-parameter order if the rescaling is made of
        per-parameter maps the device evaluates, else None."""
        return parameter_maps(self._reparameterisation, self.prime_parameters, self.model.names, self.parameters)

    def _fused_rules(self):
        """The truncation rules as a name -> rule dict if the device loop implements all of
        them (latent radius, minimum log q, likelihood threshold with a device likelihood),
        else None."""
        rules = {}
        for rule in self._truncation_scheme.rules:
            if rule.name in rules or rule.name not in ("latent_radius", "min_log_q", "likelihood_threshold"):
                return None
            rules[rule.name] = rule
        if "likelihood_threshold" in rules and not hasattr(self.model, "log_likelihood_torch"):
            return None
        return rules

    # ---------------------------------------------------------------- populate
    def populate(self, worst_point, n_samples=10000, plot=True, r=None, max_samples=1_000_000):
        """flowproposal.py:391-534; the ``while n_accepted < n_samples`` loop runs
        on the device when eligible."""
        st = datetime.datetime.now()
        maps = self._parameter_maps() if self.initialised else None
        rules = self._fused_rules() if self.initialised else None
        eligible = maps is not None and rules is not None
        if eligible and not maps.affine and len(maps.kind) > GeneralPopulateEngine.MAX_D:
            eligible = False
        # map_to_unit_hypercube (flowproposal/base.py:744-745,781-782, flowproposal.py:441-446): the loop
        # works on unit-hypercube values -- bounds [0, 1), the model's unit-hypercube prior -- and
        # convert_to_samples maps the pool to the physical space afterwards (base.py:1122-1123)
        hyper = bool(self.map_to_unit_hypercube)
        if eligible and hyper and "likelihood_threshold" in rules:
            eligible = False  # (a device likelihood takes physical parameters)
        if eligible:
            affine = maps.affine
            # auxiliary x-space parameters (the radius of a lone Angle): fields of the population
            # records that the sampler never sees (flowproposal/base.py:1100-1128)
            aux = len(maps.names) > len(self.model.names)
            engine_cls = PopulateEngine if affine else GeneralPopulateEngine
            if (self._engine is None or self._engine.flow is not self.flow
                    or type(self._engine) is not engine_cls or self._engine.names != list(maps.names)):
                # (the class can change between populates: boundary inversion re-detects its
                # edges whenever the flow is retrained, rescale.py:662-665)
                self._engine = engine_cls(
                    self.flow, maps.names, self.population_dtype,
                    row_template=nessai_empty_structured_array(1, dtype=self.population_dtype),
                )
                if self.device_prior not in ("auto", True, "uniform"):
                    self._log_prior_const = None
                elif hyper:
                    # the stock Model.log_prior_unit_hypercube is log(1) inside the cube (model.py:593-601)
                    stock = type(self.model).log_prior_unit_hypercube is _ReferenceModel.log_prior_unit_hypercube
                    self._log_prior_const = 0.0 if stock else None
                else:
                    self._log_prior_const = detect_uniform_box_prior(self.model, self.rng)
            self._engine.n_model = len(self.model.names)
            if self.accumulate_weights and self._log_prior_const is None:
                eligible = False  # the accumulating loop keeps everything on the device
        if not eligible:
            logger.debug("B200: configuration not eligible for the fused loop; using the host loop")
            return super().populate(worst_point, n_samples=n_samples, plot=plot, r=r, max_samples=max_samples)
        self._truncation_scheme.prepare(self, worst_point, radius=r)
        if self.indices:
            logger.debug("Existing pool of samples is not empty. Discarding existing samples.")
        self.indices = []
        unbounded = (-np.inf, np.inf)
        bounds = ({n: (0.0, float(np.nextafter(1.0, 0.0))) for n in self.model.names} if hyper  # (x >= 1 is outside)
                  else self.model.bounds)
        lo = [bounds.get(n, unbounded)[0] for n in maps.names]
        hi = [bounds.get(n, unbounded)[1] for n in maps.names]
        t = self.latent_temperature
        in_loop = "likelihood_threshold" in rules
        extra = {} if affine else dict(pre_scale=maps.pre_scale, pre_shift=maps.pre_shift, src=maps.src)
        self._engine.configure(
            *((maps.scale, maps.shift) if affine else (maps.kind, maps.scale, maps.shift)),
            lo, hi, self._log_prior_const,
            rules["latent_radius"].threshold if "latent_radius" in rules else 0.0,
            1.0 if t in (None, 1.0) else float(np.sqrt(t)),
            min_log_q=rules["min_log_q"].min_log_q if "min_log_q" in rules else None,
            likelihood=self.model.log_likelihood_torch if in_loop else None,
            log_l_threshold=rules["likelihood_threshold"].threshold if in_loop else None,
            **extra,
        )
        # A prior that is not a constant is evaluated on the host each turn.  The priors of auxiliary
        # parameters (the chi radius of Angle / AnglePair / ToCartesian, the augment parameters) are
        # added by the tail kernel, so with those only the MODEL's prior is asked for here
        # (flowproposal/base.py:1040-1067: log_prior = model prior + reparameterisation priors).
        if self._log_prior_const is not None:
            host_prior = None
        elif aux:
            host_prior = (self.model.batch_evaluate_log_prior_unit_hypercube if hyper
                          else self.model.batch_evaluate_log_prior)
        else:
            host_prior = self.unit_hypercube_log_prior if hyper else self.log_prior
        evals0 = getattr(self._engine, "likelihood_evaluations", 0)
        if self.accumulate_weights:
            rows, n_proposed, n_accepted = self._engine.run_accumulate(
                int(n_samples), int(self.drawsize), max_samples=max_samples
            )
        else:
            rows, n_proposed, n_accepted = self._engine.run(
                int(n_samples), int(self.drawsize), max_samples=max_samples, host_prior=host_prior
            )
        self.x = rows
        # (with auxiliary parameters the reference's own repacking drops them and sets logP)
        self.samples = self.convert_to_samples(self.x, plot=plot) if (host_prior is not None or aux or hyper) else rows
        if self._plot_pool and plot:
            self.plot_pool(self.samples)
        self.population_time += datetime.datetime.now() - st
        if in_loop:
            self.model.likelihood_evaluations += self._engine.likelihood_evaluations - evals0
        else:
            logger.debug("Evaluating log-likelihoods")
            fn = getattr(self.model, "log_likelihood_torch", None)
            device_ok = fn is not None and host_prior is None and not aux and not hyper and len(rows)
            if device_ok and self._engine.world == 1:
                # the accepted records are still on the device: 8 bytes per row come back
                self.samples["logL"] = self._engine.device_log_likelihood(len(rows), fn).cpu().numpy()
                self.model.likelihood_evaluations += len(rows)
            elif device_ok and self._engine.sharded_pool_likelihood(self.samples, fn):
                # several GPUs, pool in shared host memory: every rank evaluated its own records
                self.model.likelihood_evaluations += len(rows)
            else:
                self.samples["logL"] = self.model.batch_evaluate_log_likelihood(self.samples)
        if self.check_acceptance:
            self.acceptance.append(self.compute_acceptance(worst_point["logL"]))
        self.indices = IndexPool(self.rng.permutation(self.samples.size))
        self.population_acceptance = n_accepted / n_proposed
        self.populated_count += 1
        self.populated = True
        self._checked_population = False

    def __getstate__(self):
        state = super().__getstate__()
        state.pop("_engine", None)
        return state

    def resume(self, model, flow_config, weights_file=None):
        self._engine = None
        super().resume(model, flow_config, weights_file=weights_file)


def _check_dtype_surface():
    """The device writes live-point records with nessai's own dtype defaults."""
    assert nessai_config.livepoints.default_float_dtype == "f8"


_check_dtype_surface()


# ------------------------------------------------------------------ augmented flows
class B200AugmentedFlowProposal(B200NessaiFlowProposal, AugmentedFlowProposal):
    """``AugmentedFlowProposal`` (/root/reference/src/nessai/proposal/augmented.py:21-260) on the
    B200: the flow over the ``dims + augment_dims`` inputs (custom mask, augmented.py:91-96) is
    trained and evaluated by the CUDA kernels, and the populate loop runs on the device as for
    ``B200NessaiFlowProposal`` -- the augment parameters ``e_i`` are extra x-space parameters passed
    through unchanged whose N(0, 1) prior (augmented.py:162-178) the tail kernel adds to the log
    prior (kind 17); like the auxiliary radius of ``Angle`` they are fields of the population
    records that ``convert_to_samples`` drops.  With ``marginalise_augment=True`` the proposal
    density of every row is a Monte-Carlo marginal over ``n_marg`` forward passes (:180-200): that
    loop stays the reference's, its passes reach ``forward_and_log_prob`` as ONE batch of
    ``n * n_marg`` rows.

    A module-level class: the sampler's checkpoint pickles the proposal
    (samplers/base.py:346, utils/io.py:112), which needs a stable ``__module__.__qualname__``."""

    _FlowModelClass = B200FlowModel

    def _parameter_maps(self):
        if self.marginalise_augment or self.map_to_unit_hypercube:
            # (log q is the marginal estimate of augmented.py:180-200: the reference's loop; the reference
            # itself refuses the unit hypercube with augment parameters, :150-154)
            return None
        aug = list(self.augment_parameters)
        n = len(aug)
        if n == 0 or self.prime_parameters[-n:] != aug or self.parameters[-n:] != aug:
            return None
        base = parameter_maps(self._reparameterisation, self.prime_parameters[:-n], self.model.names,
                              self.parameters[:-n])
        if base is None:
            return None
        d0 = len(base.kind)
        src = np.concatenate([base.src, np.repeat(np.arange(d0, d0 + n, dtype=np.int32)[:, None], 3, axis=1)])
        cat = lambda a, v, dt: np.concatenate([a, np.full(n, v, dtype=dt)])  # noqa: E731
        return ParameterMaps(cat(base.kind, KIND_GAUSS_AUX, np.int32), cat(base.scale, 1.0, np.float64),
                             cat(base.shift, 0.0, np.float64), cat(base.pre_scale, 1.0, np.float64),
                             cat(base.pre_shift, 0.0, np.float64), src, list(base.names) + aug)

    def resume(self, model, flow_config, weights_file=None):
        """The pickled state keeps the custom mask as the ndarray ``__init__`` built
        (augmented.py:91-96, flowproposal/base.py:1294-1295), but the reference's ``resume`` only
        handles a list (base.py:1256-1259 leaves ``m`` unbound otherwise): hand it a list."""
        if getattr(self, "mask", None) is not None and not isinstance(self.mask, list):
            self.mask = np.asarray(self.mask).tolist()
        super().resume(model, flow_config, weights_file=weights_file)


# ------------------------------------------------------------------ importance nested sampler
def __getattr__(name):
    # lazily re-exported: importing the importance sampler pulls in more of nessai
    if name in ("B200ImportanceFlowProposal", "B200ImportanceNestedSampler"):
        from . import nessai_ins_plugin

        return getattr(nessai_ins_plugin, name)
    raise AttributeError(name)
