"""Bindings into the real ``nessai`` package (plugin points P2 / P3).

Importing this module requires ``nessai`` to be importable.  It defines

* ``B200NessaiFlowProposal(nessai.proposal.FlowProposal)`` with
  ``_FlowModelClass = B200FlowModel`` -- every flow evaluation of the unmodified
  reference proposal (train / forward_pass / backward_pass / truncation rules)
  runs on the CUDA kernels -- and a ``populate`` override that runs the fused
  device loop whenever the configuration allows it (z-score / null
  reparameterisation, latent-radius truncation only, no weight accumulation) and
  otherwise defers to the reference's host loop;
* the entry point ``nessai.proposals: b200flowproposal``, so
  ``FlowSampler(model, flow_proposal_class="b200flowproposal")`` picks it up
  through ``nessai.proposal.utils.get_flow_proposal_class``
  (/root/reference/src/nessai/proposal/utils.py:110-155).
"""

from __future__ import annotations

import datetime
import logging

import numpy as np
from nessai import config as nessai_config
from nessai.livepoint import empty_structured_array as nessai_empty_structured_array
from nessai.proposal.flowproposal import FlowProposal
from nessai.reparameterisations import NullReparameterisation, RescaleToBounds, ScaleAndShift

from .flowmodel import B200FlowModel
from .proposal import IndexPool, PopulateEngine, detect_uniform_box_prior

logger = logging.getLogger(__name__)



def diagonal_rescaling(rep, prime_parameters, model_names):
    """``(scale, shift)`` with ``x = x' * scale + shift`` per parameter (model order) when the
    combined reparameterisation ``rep`` (reparameterisations/combined.py:154-192) is a diagonal
    affine map, else ``None``.  A diagonal affine is what the fused populate tail evaluates
    (float64): ``log|J| = sum log|scale|``.  Recognised:

    * ``NullReparameterisation`` (reparameterisations/null.py): identity;
    * ``ScaleAndShift`` / ``Rescale`` (rescale.py:233-291) without pre-/post-rescaling:
      ``x = x' * scale + shift``;
    * ``RescaleToBounds`` (rescale.py:321-731) without boundary inversion and without pre-/post-
      rescaling functions: ``x = (hi - lo) * (x' - r0) / (r1 - r0) + lo + offset`` with
      ``[lo, hi]`` the current (possibly data-updated) bounds, ``[r0, r1]`` the rescale bounds
      (rescale.py:533-543,669-680), i.e. ``scale = (hi - lo) / (r1 - r0)``,
      ``shift = lo + offset - scale * r0``.

    Anything else (logit / log post-rescaling, boundary inversion, angles, user classes) is not
    a diagonal affine and keeps the reference's host loop."""
    if rep is None:
        return None
    if list(prime_parameters) != [p for r in rep.values() for p in r.output_parameters]:
        return None
    scale, shift, names = [], [], []
    for r in rep.values():
        if isinstance(r, ScaleAndShift):
            if r.has_pre_rescaling or r.has_post_rescaling or not r.one_to_one:
                return None
            for p in r.parameters:
                scale.append(float(r.scale[p]))
                shift.append(float(r.shift[p]) if r.shift else 0.0)
                names.append(p)
        elif isinstance(r, NullReparameterisation):
            for p in r.parameters:
                scale.append(1.0)
                shift.append(0.0)
                names.append(p)
        elif type(r) is RescaleToBounds:
            # subclasses may override the rescaling hooks: only the stock class is recognised
            if r.boundary_inversion or r.has_pre_rescaling or r.has_post_rescaling or not r.one_to_one:
                return None
            for p in r.parameters:
                lo, hi = (float(b) for b in r.bounds[p])
                s = (hi - lo) / float(r._rescale_factor[p])
                scale.append(s)
                shift.append(lo + float(r.offsets[p]) - s * float(r._rescale_shift[p]))
                names.append(p)
        else:
            return None
    if names != list(model_names):
        return None
    scale, shift = np.asarray(scale, dtype=np.float64), np.asarray(shift, dtype=np.float64)
    if not (np.all(np.isfinite(scale)) and np.all(np.isfinite(shift)) and np.all(scale != 0.0)):
        return None
    return scale, shift


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
    def _diagonal_rescaling(self):
        """(scale, shift) in prime-parameter order if the rescaling is a diagonal
        affine handled on the device, else None."""
        if self.map_to_unit_hypercube or self.accumulate_weights:
            return None
        return diagonal_rescaling(self._reparameterisation, self.prime_parameters, self.model.names)

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
        diag = self._diagonal_rescaling() if self.initialised else None
        rules = self._fused_rules() if self.initialised else None
        if diag is None or rules is None:
            logger.debug("B200: configuration not eligible for the fused loop; using the host loop")
            return super().populate(worst_point, n_samples=n_samples, plot=plot, r=r, max_samples=max_samples)
        st = datetime.datetime.now()
        if not self.initialised:
            raise RuntimeError(
                "Proposal has not been initialised. Try calling `initialise()` first."
            )
        self._truncation_scheme.prepare(self, worst_point, radius=r)
        if self.indices:
            logger.debug("Existing pool of samples is not empty. Discarding existing samples.")
        self.indices = []
        if self._engine is None or self._engine.flow is not self.flow:
            self._engine = PopulateEngine(
                self.flow, self.model.names, self.population_dtype,
                row_template=nessai_empty_structured_array(1, dtype=self.population_dtype),
            )
            self._log_prior_const = (
                detect_uniform_box_prior(self.model, self.rng)
                if self.device_prior in ("auto", True, "uniform")
                else None
            )
        lo = [self.model.bounds[n][0] for n in self.model.names]
        hi = [self.model.bounds[n][1] for n in self.model.names]
        t = self.latent_temperature
        in_loop = "likelihood_threshold" in rules
        self._engine.configure(
            diag[0], diag[1], lo, hi, self._log_prior_const,
            rules["latent_radius"].threshold if "latent_radius" in rules else 0.0,
            1.0 if t in (None, 1.0) else float(np.sqrt(t)),
            min_log_q=rules["min_log_q"].min_log_q if "min_log_q" in rules else None,
            likelihood=self.model.log_likelihood_torch if in_loop else None,
            log_l_threshold=rules["likelihood_threshold"].threshold if in_loop else None,
        )
        host_prior = None if self._log_prior_const is not None else self.log_prior
        evals0 = getattr(self._engine, "likelihood_evaluations", 0)
        rows, n_proposed, n_accepted = self._engine.run(
            int(n_samples), int(self.drawsize), max_samples=max_samples, host_prior=host_prior
        )
        self.x = rows
        self.samples = self.convert_to_samples(self.x, plot=plot) if host_prior is not None else rows
        if self._plot_pool and plot:
            self.plot_pool(self.samples)
        self.population_time += datetime.datetime.now() - st
        if in_loop:
            self.model.likelihood_evaluations += self._engine.likelihood_evaluations - evals0
        else:
            logger.debug("Evaluating log-likelihoods")
            fn = getattr(self.model, "log_likelihood_torch", None)
            if fn is not None and self._engine.world == 1 and host_prior is None and len(rows):
                # the accepted records are still on the device: 8 bytes per row come back
                self.samples["logL"] = self._engine.device_log_likelihood(len(rows), fn).cpu().numpy()
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


# ------------------------------------------------------------------ importance nested sampler
def __getattr__(name):
    # lazily re-exported: importing the importance sampler pulls in more of nessai
    if name in ("B200ImportanceFlowProposal", "B200ImportanceNestedSampler"):
        from . import nessai_ins_plugin

        return getattr(nessai_ins_plugin, name)
    raise AttributeError(name)
