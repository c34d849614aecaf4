"""Host logic of the populate loops on a SIMULATED device (tests/_simdevice.py: the C-ABI entry
points answered by the float64 oracle through the same raw pointers).  Covers what a GPU is not
needed for: slot offsets and Philox counter bookkeeping of ``run_accumulate``, the turn sequence
of ``GeneralPopulateEngine`` (draw -> tail -> float64-row rejection step), argument order of
every call, record layout, and the engine selection / configuration of the nessai plugin."""

import numpy as np
import pytest
from conftest import load_golden, reference_or_skip

import _simdevice


def _flow():
    from oracle.flow_numpy import NumpyFlow

    g, cfg, sd = load_golden("c2_realnvp_mlp")
    return NumpyFlow(sd, ftype="realnvp", net="mlp", hidden_features=cfg["n_neurons"]), cfg["n_inputs"]


def _zscore(D):
    live = 1.5 * np.random.default_rng(5).standard_normal((500, D)) + 0.3
    return live.std(0), live.mean(0)


@pytest.mark.parametrize("n_samples,max_samples", [(150, 10**6), (10**5, 2999)])
def test_run_accumulate_matches_oracle_loop(monkeypatch, n_samples, max_samples):
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import PopulateEngine
    from oracle.philox_numpy import accept_uniform, latent_normals
    from oracle.populate_numpy import populate_loop_accumulate

    sim = _simdevice.install(monkeypatch)
    nf, D = _flow()
    names = [f"x{i}" for i in range(D)]
    eng = PopulateEngine(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
    scale, shift = _zscore(D)
    lpc, radius, drawsize = -D * np.log(8.0), 4.9, 1000
    eng.configure(scale, shift, np.full(D, -4.0), np.full(D, 4.0), lpc, radius)
    eng.seed = 4242
    eng._turn_rows = 12345  # an engine that has populated before
    rows, n_proposed, n_accepted = eng.run_accumulate(n_samples, drawsize, max_samples=max_samples)
    info = eng.last_accumulate
    assert info["stride"] == 1024 and info["draw_offsets"][0] == 12345
    # ---- the oracle loop on the same Philox rows; its uniforms are the slot rows' own counters
    state = dict(turn=0, masks=[], reject=0)

    def draw_z(n):
        z = latent_normals(eng.seed, info["draw_offsets"][state["turn"]] + np.arange(n), D)
        state["turn"] += 1
        return z

    def on_valid(valid):
        state["masks"].append(valid)

    def draw_u(m):
        base, nrows = info["rejects"][state["reject"]]
        state["reject"] += 1
        assert nrows == len(state["masks"]) * info["stride"]
        slot_rows = np.concatenate([t * info["stride"] + np.flatnonzero(v) for t, v in enumerate(state["masks"])])
        assert len(slot_rows) == m
        return accept_uniform(eng.seed, base + slot_rows)

    x, p, a, log_n_exp = populate_loop_accumulate(
        nf, draw_z, draw_u, n_samples, drawsize, max_samples=max_samples, on_valid=on_valid,
        scale=scale, shift=shift, lo=-4.0, hi=4.0, log_prior_const=lpc, r_max=radius)
    assert (n_proposed, n_accepted) == (p, a)
    assert state["turn"] == len(info["draw_offsets"]) and state["reject"] == len(info["rejects"])
    got = np.stack([rows[nm] for nm in names], axis=-1)
    assert got.shape == x.shape and len(rows) == min(a, n_samples)
    np.testing.assert_allclose(got, x, rtol=1e-6, atol=1e-6)  # x' crosses the C ABI as fp32
    assert np.all(rows["logP"] == lpc)
    np.testing.assert_allclose(np.log(info["n_expected"][-1]), log_n_exp, rtol=1e-12)
    # ---- counters: every draw and every rejection step has its own block, in order
    spans = [(o, o + drawsize) for o in info["draw_offsets"]] + [(b, b + r) for b, r in info["rejects"]]
    spans.sort()
    assert all(s[1] <= t[0] for s, t in zip(spans, spans[1:])) and eng._turn_rows == spans[-1][1]
    if max_samples == 2999:
        assert len(info["draw_offsets"]) == 3 and len(rows) < n_samples  # ended by max_samples
    else:
        assert len(info["draw_offsets"]) >= 2 and len(rows) == n_samples
    # the engine is still usable for the ordinary loop afterwards (buffers, counters)
    rows2, p2, a2 = eng.run(50, drawsize, max_samples=10**6)
    assert len(rows2) == 50 and a2 >= 50
    assert [c[0] for c in sim.calls].count("sum_exp") == len(info["draw_offsets"])


def test_general_engine_loop_matches_oracle(monkeypatch):
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import GeneralPopulateEngine
    from oracle.philox_numpy import accept_uniform, latent_normals
    from oracle.reparam_numpy import tail_rows

    sim = _simdevice.install(monkeypatch)
    nf, D = _flow()
    names = [f"x{i}" for i in range(D)]
    eng = GeneralPopulateEngine(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
    kind = (np.arange(D) % 4).astype(np.int32)
    scale = np.where(kind == 1, 8.0, np.where(kind == 3, 0.5, np.where(kind == 2, -2.0, 1.4)))
    shift = np.where(kind == 1, -4.0, np.where(kind == 2, 3.0, 0.1))
    lo = np.where(kind == 1, -4.0, np.where(kind == 2, -3.0, np.where(kind == 3, 0.0, -5.0)))
    hi = np.where(kind == 1, 4.0, np.where(kind == 2, 3.0, np.where(kind == 3, 6.0, 5.0)))
    lpc, r_max, mlq, n = -3.0, 4.9, -27.0, 2000
    eng.configure(kind, scale, shift, lo, hi, lpc, r_max, 1.0, min_log_q=mlq)
    eng.seed = 77
    rows, n_proposed, n_accepted = eng.run(300, n, max_samples=10**6)
    turns = n_proposed // n
    assert [c[0] for c in sim.calls] == ["draw", "tail", "accept_x64"] * turns and turns >= 2
    # oracle: same rows of the Philox stream, turn by turn
    kept, total = [], 0
    for t in range(turns):
        z = latent_normals(77, t * n + np.arange(n), D)
        with np.errstate(all="ignore"):
            xp, lq_flow = nf.sample_and_log_prob(z)
        keep = np.sqrt(np.sum(z**2, axis=1)) <= r_max
        x, lq, lw, valid = tail_rows(xp.astype(np.float32), np.where(keep, lq_flow, np.nan), kind=kind, scale=scale,
                                     shift=shift, lo=lo, hi=hi, log_prior_const=lpc, min_log_q=mlq)
        u = accept_uniform(77, t * n + np.arange(n))
        acc = valid & ((lw - lw[valid].max()) > np.log(u))
        kept.append(x[acc][: max(300 - total, 0)])
        total += int(acc.sum())
    assert total == n_accepted and len(rows) == 300
    got = np.stack([rows[nm] for nm in names], axis=-1)
    np.testing.assert_allclose(got, np.concatenate(kept), rtol=1e-12, atol=1e-12)
    assert np.all((got >= lo) & (got <= hi)) and np.all(rows["logP"] == lpc)


def test_general_engine_accumulate(monkeypatch):
    """accumulate_weights over the non-affine tail: with identity maps it is the affine engine's
    accumulating loop row for row; with real maps the pool follows the accumulated weights."""
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import GeneralPopulateEngine, PopulateEngine
    from oracle.philox_numpy import accept_uniform

    sim = _simdevice.install(monkeypatch)
    nf, D = _flow()
    names = [f"x{i}" for i in range(D)]
    scale, shift = _zscore(D)
    lo, hi, lpc, radius, drawsize = np.full(D, -4.0), np.full(D, 4.0), -D * np.log(8.0), 4.9, 1000
    aff = PopulateEngine(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
    aff.configure(scale, shift, lo, hi, lpc, radius, min_log_q=-40.0)
    gen = GeneralPopulateEngine(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
    gen.configure(np.zeros(D, dtype=np.int32), scale, shift, lo, hi, lpc, radius, min_log_q=-40.0)
    aff.seed = gen.seed = 31
    ra, pa, aa = aff.run_accumulate(150, drawsize, max_samples=10**6)
    rg, pg, ag = gen.run_accumulate(150, drawsize, max_samples=10**6)
    assert (pa, aa) == (pg, ag) and aff._turn_rows == gen._turn_rows
    assert aff.last_accumulate["rejects"] == gen.last_accumulate["rejects"]
    np.testing.assert_allclose(aff.last_accumulate["n_expected"], gen.last_accumulate["n_expected"], rtol=1e-12)
    for nm in names:  # (the affine path forms x from fp32 x' at the rejection step, the tail before it)
        np.testing.assert_allclose(rg[nm], ra[nm], rtol=1e-12, atol=1e-12)
    # sigmoid / abs / exp maps: the final rejection step over all slots, replayed from the buffers
    kind = (np.arange(D) % 4).astype(np.int32)
    sc = np.where(kind == 1, 8.0, np.where(kind == 3, 0.5, np.where(kind == 2, -2.0, 1.4)))
    sh = np.where(kind == 1, -4.0, np.where(kind == 2, 3.0, 0.1))
    lo2 = np.where(kind == 1, -4.0, np.where(kind == 2, -3.0, np.where(kind == 3, 0.0, -5.0)))
    hi2 = np.where(kind == 1, 4.0, np.where(kind == 2, 3.0, np.where(kind == 3, 6.0, 5.0)))
    gen.configure(kind, sc, sh, lo2, hi2, -3.0, radius, min_log_q=-27.0)
    rows, p, a = gen.run_accumulate(200, drawsize, max_samples=10**6)
    info = gen.last_accumulate
    turns, stride = len(info["draw_offsets"]), info["stride"]
    lw = gen.d_logw[: turns * stride].numpy()
    x64 = gen.d_x64[: turns * stride].numpy()
    stats = gen.d_stats.numpy()
    ok = ~np.isnan(lw)
    assert stats[1] == ok.sum() and stats[0] == lw[ok].max() and len(rows) == 200 and a >= 200
    base, nrows = info["rejects"][-1]
    acc = ok & ((lw - stats[0]) > np.log(accept_uniform(gen.seed, base + np.arange(nrows))))
    assert a == acc.sum()
    got = np.stack([rows[nm] for nm in names], axis=-1)
    np.testing.assert_array_equal(got, x64[acc][:200])
    assert np.all((got >= lo2) & (got <= hi2))
    assert [c[0] for c in sim.calls if c[0] in ("draw", "tail")][-2 * turns:] == ["draw", "tail"] * turns


@pytest.mark.reference
@pytest.mark.parametrize("variant", ["zscore", "rescaletobounds", "logit_mixed", "inversion_edges", "accumulate",
                                     "accumulate_min_log_q", "likelihood_threshold", "logit_likelihood_threshold",
                                     "zscore_gaussian_cdf", "angle_aux", "angle_and_radial_parameter",
                                     "accumulate_logit", "accumulate_likelihood_threshold",
                                     "angle_aux_likelihood_threshold", "accumulate_angle_likelihood_threshold",
                                     "to_cartesian", "angle_pair_aux", "dequantise", "unit_hypercube",
                                     "unit_hypercube_logit", "augmented", "augmented_logit",
                                     "angle_aux_host_prior", "augmented_host_prior", "dequantise_logit"])
def test_plugin_populate_on_simulated_device(tmp_path, monkeypatch, variant):
    """``B200NessaiFlowProposal.populate`` end to end with the reference's own proposal object
    (reparameterisations, truncation scheme, live-point dtype): engine selection, configuration
    and the pool it hands to the sampler, against the reference's host ``populate`` of the same
    trained flow (distributional: the random streams differ)."""
    reference_or_skip()
    import torch
    from nessai.flowmodel import FlowModel
    from nessai.livepoint import numpy_array_to_live_points
    from nessai.model import Model

    from nessai_b200.nessai_plugin import B200AugmentedFlowProposal, B200NessaiFlowProposal
    from nessai_b200.proposal import GeneralPopulateEngine, PopulateEngine
    from oracle.flow_numpy import NumpyFlow

    D = 4
    names = [f"x{i}" for i in range(D)]

    class Box(Model):
        def __init__(self):
            self.names = list(names)
            self.bounds = {n: [-5.0, 5.0] for n in names}
            if "angle" in variant:  # x0: an angle in [0, 2 pi]; x1: a radius-like parameter
                self.bounds["x0"] = [0.0, 2 * np.pi]
                self.bounds["x1"] = [0.0, np.pi] if variant == "angle_pair_aux" else [0.0, 5.0]  # (a zenith angle)
            if variant.startswith("dequantise"):  # x0: a discrete parameter 0 .. 4
                self.bounds["x0"] = [0.0, 4.0]

        def new_point(self, N=1):
            x = super().new_point(N)
            if variant.startswith("dequantise"):
                x["x0"] = np.floor(x["x0"])
            return x

        def to_unit_hypercube(self, x):  # map_to_unit_hypercube (model.py:603-614: the user supplies both maps)
            u = x.copy()
            for n in self.names:
                u[n] = (x[n] - self.bounds[n][0]) / (self.bounds[n][1] - self.bounds[n][0])
            return u

        def from_unit_hypercube(self, u):
            x = u.copy()
            for n in self.names:
                x[n] = u[n] * (self.bounds[n][1] - self.bounds[n][0]) + self.bounds[n][0]
            return x

        def log_prior(self, x):
            lp = np.log(self.in_bounds(x), dtype="float") + LOG_P
            if "host_prior" in variant:  # not a constant: evaluated on the host every turn
                lp = lp - 0.08 * (x["x2"] - 1.0) ** 2
            return lp

        def log_likelihood(self, x):
            a = self.unstructured_view(x)
            if "angle" in variant:  # periodic in the angle
                return np.cos(a[..., 0] - 1.0) - 0.5 * np.sum((a[..., 1:] - 1.0) ** 2, axis=-1)
            return -0.5 * np.sum(a**2, axis=-1)

        def log_likelihood_torch(self, x):  # INTEGRATION.md 3a (host tensors on the simulated device)
            self.device_rows = getattr(self, "device_rows", 0) + x.shape[0]
            if "angle" in variant:
                return torch.cos(x[:, 0] - 1.0) - 0.5 * ((x[:, 1:] - 1.0) ** 2).sum(dim=1)
            return -0.5 * (x * x).sum(dim=1)

    class CpuFlowB200Proposal(B200AugmentedFlowProposal if variant.startswith("augmented") else B200NessaiFlowProposal):
        _FlowModelClass = FlowModel  # the reference's CPU flow; the simulated device evaluates its weights

    kw = dict(
        zscore={}, rescaletobounds=dict(fallback_reparameterisation="rescaletobounds"),
        logit_mixed=dict(reparameterisations={"x0": "logit", "x1": "default", "x2": "z-score", "x3": "log-rescale"}),
        inversion_edges=dict(reparameterisations={"inversion": dict(parameters=names)}),
        accumulate=dict(accumulate_weights=True),
        accumulate_min_log_q=dict(accumulate_weights=True, truncation_methods=["latent_radius", "min_log_q"]),
        accumulate_likelihood_threshold=dict(accumulate_weights=True,
                                             truncation_methods=["latent_radius", "likelihood_threshold"]),
        angle_aux_likelihood_threshold=dict(
            truncation_methods=["latent_radius", "likelihood_threshold"],
            reparameterisations={"x0": "angle", "x1": "default", "x2": "z-score", "x3": "logit"}),
        accumulate_angle_likelihood_threshold=dict(
            accumulate_weights=True, truncation_methods=["latent_radius", "likelihood_threshold"],
            reparameterisations={"x0": "angle", "x1": "default", "x2": "z-score", "x3": "logit"}),
        accumulate_logit=dict(accumulate_weights=True,
                              reparameterisations={"x0": "logit", "x1": "default", "x2": "default", "x3": "logit"}),
        zscore_gaussian_cdf=dict(reparameterisations={"zscore-gaussian-cdf": dict(parameters=names)}),
        angle_aux=dict(reparameterisations={"x0": "angle", "x1": "default", "x2": "z-score", "x3": "logit"}),
        angle_and_radial_parameter=dict(reparameterisations={"angle": {"parameters": ["x0", "x1"]},
                                                             "x2": "default", "x3": "default"}),
        angle_aux_host_prior=dict(reparameterisations={"x0": "angle", "x1": "default", "x2": "z-score", "x3": "logit"}),
        augmented_host_prior=dict(augment_dims=1),
        augmented=dict(augment_dims=2),  # proposal/augmented.py: two augment parameters e_0, e_1
        augmented_logit=dict(augment_dims=1, reparameterisations={"x0": "logit", "x1": "default", "x2": "z-score",
                                                                  "x3": "logit"}),
        unit_hypercube=dict(map_to_unit_hypercube=True),
        unit_hypercube_logit=dict(map_to_unit_hypercube=True,
                                  reparameterisations={"x0": "logit", "x1": "default", "x2": "z-score", "x3": "logit"}),
        to_cartesian=dict(reparameterisations={"x0": "to-cartesian", "x1": "default", "x2": "z-score", "x3": "logit"}),
        angle_pair_aux=dict(reparameterisations={"angle-pair": {"parameters": ["x0", "x1"]}, "x2": "default",
                                                 "x3": "logit"}),
        dequantise=dict(reparameterisations={"x0": "dequantise", "x1": "default", "x2": "z-score", "x3": "logit"}),
        dequantise_logit=dict(reparameterisations={"x0": "dequantise-logit", "x1": "default", "x2": "z-score",
                                                   "x3": "logit"}),
        likelihood_threshold=dict(truncation_methods=["latent_radius", "likelihood_threshold"]),
        logit_likelihood_threshold=dict(truncation_methods=["latent_radius", "likelihood_threshold"],
                                        reparameterisations={"x0": "logit", "x1": "logit", "x2": "default",
                                                             "x3": "default"}),
    )[variant]
    contour = variant.endswith("likelihood_threshold")
    LOG_P = -D * np.log(10.0) if "angle" not in variant else -np.log(2 * np.pi * 5.0 * 100.0)
    if variant == "angle_pair_aux":
        LOG_P = -np.log(2 * np.pi * np.pi * 100.0)
    if variant.startswith("dequantise"):
        LOG_P = -np.log(4.0 * 1000.0)
    model = Box()
    rng = np.random.default_rng(9)
    model.set_rng(rng)
    torch.manual_seed(9)
    flow_config = dict(n_blocks=2, n_neurons=8, n_layers=1, net="mlp", batch_norm_between_layers=False)
    common = dict(rng=rng, flow_config=flow_config, training_config=dict(max_epochs=30, patience=30),
                  output=str(tmp_path), poolsize=400, drawsize=2000, plot=False)
    prop = CpuFlowB200Proposal(model, **common, **kw)
    prop.initialise()
    pts = np.clip(1.2 * rng.standard_normal((600, D)) + 0.5, -4.9, 4.9)
    if "angle" in variant:
        pts[:, 0] = (1.0 + 0.8 * rng.standard_normal(600)) % (2 * np.pi)
        pts[:, 1] = np.clip(np.abs(1.0 + 0.7 * rng.standard_normal(600)), 0.05, 3.0 if variant == "angle_pair_aux" else 4.9)
    if variant.startswith("dequantise"):
        pts[:, 0] = rng.integers(0, 5, 600)
    live = numpy_array_to_live_points(pts, names)
    live["logL"] = model.log_likelihood(live)
    prop.train(live, plot=False)
    if variant == "inversion_edges":
        (r,) = prop._reparameterisation.values()
        r._edges.update(x0="lower", x1="upper", x2=False, x3="lower")
    # the reference's own host populate of the same flow, for comparison
    worst = live[np.argsort(live["logL"])[len(live) // 2 if contour else 0]]  # a contour that cuts the pool
    from nessai.proposal.flowproposal import FlowProposal

    FlowProposal.populate(prop, worst, n_samples=400, plot=False)
    ref = np.stack([prop.samples[n] for n in names], axis=-1).copy()
    ref_acceptance, ref_dtype, ref_x_dtype = prop.population_acceptance, prop.samples.dtype, prop.x.dtype
    # the simulated device evaluates the trained weights with the float64 oracle
    sim = _simdevice.install(monkeypatch)
    sd = {k: v.detach().cpu().numpy() for k, v in prop.flow.model.state_dict().items()}
    nf = NumpyFlow(sd, ftype="realnvp", net="mlp", hidden_features=8)
    prop.flow.model._ready = lambda: None
    prop.flow.model._handle = _simdevice.SimHandle(nf, len(prop.prime_parameters))  # D + auxiliary parameters
    prop.populate(worst, n_samples=400, plot=False)
    assert prop._engine is not None and len(sim.calls) > 0  # not the host loop
    general = variant in ("logit_mixed", "inversion_edges", "logit_likelihood_threshold", "zscore_gaussian_cdf",
                          "angle_aux", "angle_and_radial_parameter", "accumulate_logit",
                          "angle_aux_likelihood_threshold", "accumulate_angle_likelihood_threshold",
                          "to_cartesian", "angle_pair_aux", "dequantise", "unit_hypercube_logit", "augmented",
                          "augmented_logit", "angle_aux_host_prior", "augmented_host_prior", "dequantise_logit")
    if variant.startswith("augmented"):  # the augment parameters never reach the sampler (augmented.py, base.py:1100-1128)
        aug = [f"e_{i}" for i in range(prop.augment_dims)]
        assert prop._engine.names == names + aug and prop.samples.dtype.names[:D] == tuple(names)
        assert all(a in prop.x.dtype.names and a not in prop.samples.dtype.names for a in aug)
    if variant.startswith("unit_hypercube"):  # the loop ran on unit-hypercube values, the sampler gets physical ones
        u = np.stack([prop.x[n] for n in names], axis=-1)
        assert np.all((u >= 0.0) & (u < 1.0)) and prop.x.size == prop.samples.size
        np.testing.assert_allclose(np.stack([prop.samples[n] for n in names], axis=-1), 10.0 * u - 5.0, rtol=0, atol=1e-12)
    if variant in ("to_cartesian", "angle_pair_aux"):
        aux = "x0_radial" if variant == "to_cartesian" else "x0_x1_radial"
        assert prop._engine.names == names + [aux] and aux in prop.x.dtype.names and aux not in prop.samples.dtype.names
    if variant.startswith("dequantise"):
        assert set(np.unique(prop.samples["x0"])) <= {0.0, 1.0, 2.0, 3.0, 4.0}
    if "angle_aux" in variant or variant == "accumulate_angle_likelihood_threshold":  # the auxiliary radius never reaches the sampler (flowproposal/base.py:1100-1128)
        assert prop._engine.names == names + ["x0_radial"] and prop.samples.dtype.names[:D] == tuple(names)
        assert "x0_radial" in prop.x.dtype.names and "x0_radial" not in prop.samples.dtype.names
    if contour:  # the likelihood ran on the "device", inside the loop, never on the host
        assert model.device_rows > 0 and np.all(prop.samples["logL"] > worst["logL"])
        assert np.all(prop.samples["logL"] > worst["logL"])
    assert type(prop._engine) is (GeneralPopulateEngine if general else PopulateEngine)
    kinds = {c[0] for c in sim.calls}
    assert ("tail" in kinds) == general and ("sum_exp" in kinds) == variant.startswith("accumulate")
    got = np.stack([prop.samples[n] for n in names], axis=-1)
    assert prop.populated and len(prop.indices) == prop.samples.size
    if variant.startswith("accumulate"):
        assert 0 < len(got) <= 400  # samples[accept][:n_samples]
    else:
        assert len(got) == 400
    assert prop.samples.dtype == ref_dtype and prop.x.dtype == ref_x_dtype == prop.population_dtype
    assert np.all((got >= -5) & (got <= 2 * np.pi)) and np.all(np.isfinite(prop.samples["logL"]))
    np.testing.assert_allclose(prop.samples["logP"], model.log_prior(prop.samples) if "host_prior" in variant else LOG_P)
    if "host_prior" in variant:
        assert prop._log_prior_const is None  # the model's prior really was evaluated on the host
    np.testing.assert_allclose(prop.samples["logL"], model.log_likelihood(prop.samples))
    if variant == "zscore_gaussian_cdf":
        # The flow proposes x' outside (0, 1), where the quantile function is NaN.  The reference keeps
        # such rows (NaN passes its bounds check, model.py:497-518), `log_w.max()` is then NaN
        # (flowproposal.py:492) and NOTHING is ever accepted: its pool is empty.  The device tail drops
        # rows with a non-finite log q, so the loop works; there is no reference pool to compare with.
        assert len(ref) == 0 and len(got) == 400
    else:
        # same target: pool moments and acceptance agree with the reference's host loop
        se = ref.std(0) * np.sqrt(1 / len(ref) + 1 / len(got))
        assert np.all(np.abs(got.mean(0) - ref.mean(0)) < 6 * se), (got.mean(0), ref.mean(0))
        assert np.all(np.abs(np.log(got.std(0) / ref.std(0))) < 0.35)
        # (the acceptance is set by the largest weight of each turn: heavy-tailed, hence the wide band)
        assert 0.2 < prop.population_acceptance / ref_acceptance < 5.0
    # the engine (device buffers, cached argument lists) is reused by the next populate
    engine = prop._engine
    prop.populate(worst, n_samples=100, plot=False)
    assert prop._engine is engine
    if variant in ("zscore", "logit_mixed", "angle_aux", "accumulate_logit", "likelihood_threshold"):
        # ... and the same populate with the entry points answered by the product's CUDA sources (SIMT
        # shim) instead of the oracle: same Philox rows in, the same pool out
        import _simtdevice
        from nessai_b200 import _lib
        from nessai_b200.spec import FlowSpec

        libs = _simtdevice.build(tmp_path)
        if libs is not None:
            engine.seed, engine._turn_rows = 777, 0
            engine._draw_key = engine._accept_key = None  # (the cached argument lists hold the seed)
            prop.populate(worst, n_samples=150, plot=False)
            pool_a, acc_a = np.stack([prop.samples[n] for n in names], axis=-1).copy(), prop.population_acceptance
            spec = FlowSpec(dict(flow_config, n_inputs=len(prop.prime_parameters)))
            theta = np.zeros(spec.n_theta, np.float32)
            ints = {}
            spec.load_state_dict_numpy(sd, theta, ints)
            prop.flow.model._handle = _simtdevice.SimtHandle(spec, spec.fold(theta, ints).program(True))
            simt = _simtdevice.SimtLib(libs)
            monkeypatch.setattr(_lib, "load", lambda: simt)
            engine._draw_key = engine._accept_key = None  # (cached argument lists hold the old handle)
            engine._turn_rows = 0
            prop.populate(worst, n_samples=150, plot=False)
            assert prop._engine is engine and len(simt.calls) > 0
            pool_b = np.stack([prop.samples[n] for n in names], axis=-1)
            # (a row at rounding distance from a threshold may flip between the fp32 kernels and the
            # float64 oracle; with these seeds none does)
            assert prop.population_acceptance == acc_a
            np.testing.assert_allclose(pool_b, pool_a, rtol=5e-4, atol=5e-4)
    # and the proposal still serves the sampler
    new = prop.draw(worst)
    assert new.dtype == ref_dtype and len(prop.indices) == prop.samples.size - 1


# ------------------------------------------------------------------ two ranks over gloo
def _rank_worker(rank, world, port, out):
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import _simdevice
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import GeneralPopulateEngine, PopulateEngine

    pipelined = PopulateEngine.run  # before install() replaces it by the serial loop
    _simdevice.install()
    _simdevice.install_fake_cuda_async()
    nf, D = _flow()
    names = [f"x{i}" for i in range(D)]
    scale, shift = _zscore(D)
    lpc, radius, drawsize = -D * np.log(8.0), 4.9, 1001  # ragged shards: 501 + 500
    res = {}
    # (0) the pipelined loop; with two ranks the pool is assembled in node-local shared host
    # memory (hostpool.py), every rank copying only its own records, turn-major
    eng = PopulateEngine(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
    eng.configure(scale, shift, np.full(D, -4.0), np.full(D, 4.0), lpc, radius)
    eng.seed = 4242
    pools = []
    for n_samples, max_samples in ((10**5, 3 * drawsize - 1), (40, 10**6), (75, 10**6)):
        rows, p, a = pipelined(eng, n_samples, drawsize, max_samples=max_samples)
        pools.append((np.stack([rows[nm] for nm in names], axis=-1).copy(), p, a))
    res["pipelined"] = pools
    res["shared_pool"] = getattr(eng, "_pool", None) is not None
    eng = PopulateEngine(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
    eng.configure(scale, shift, np.full(D, -4.0), np.full(D, 4.0), lpc, radius)
    eng.seed = 4242
    # (1) the ordinary loop, ended by max_samples so that every accepted row is in the pool
    rows, p, a = eng.run(10**5, drawsize, max_samples=3 * drawsize - 1)
    res["loop"] = (np.stack([rows[nm] for nm in names], axis=-1), p, a)
    # (2) accumulate_weights
    rows, p, a = eng.run_accumulate(120, drawsize, max_samples=10**6)
    res["acc"] = (np.stack([rows[nm] for nm in names], axis=-1), p, a, list(eng.last_accumulate["n_expected"]),
                  eng.last_accumulate["stride"], list(eng.last_accumulate["rejects"]))
    # (3) the non-affine tail
    gen = GeneralPopulateEngine(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
    kind = (np.arange(D) % 2).astype(np.int32)  # identity / sigmoid
    gen.configure(kind, np.where(kind == 1, 8.0, 1.4), np.where(kind == 1, -4.0, 0.1), np.full(D, -4.5),
                  np.full(D, 4.5), lpc, radius, 1.0, min_log_q=-40.0)
    gen.seed = 99
    rows, p, a = gen.run(10**5, drawsize, max_samples=3 * drawsize - 1)
    res["tail"] = (np.stack([rows[nm] for nm in names], axis=-1), p, a)
    torch.save(res, f"{out}.{world}.{rank}")
    if world > 1:
        dist.destroy_process_group()


def test_two_ranks_reproduce_one_rank(tmp_path):
    """Sharded over two ranks (gloo, simulated device) the loops propose and accept exactly the
    rows one rank does -- the Philox counter is the global row index and the normaliser is
    all-reduced -- and every rank ends up with the same pool."""
    import torch
    import torch.multiprocessing as mp
    from test_dist_gloo import _free_port

    out = str(tmp_path / "res")
    mp.spawn(_rank_worker, args=(1, _free_port(), out), nprocs=1, join=True)
    mp.spawn(_rank_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    one = torch.load(f"{out}.1.0", weights_only=False)
    r0, r1 = (torch.load(f"{out}.2.{r}", weights_only=False) for r in range(2))

    def same_rows(a, b):
        assert a.shape == b.shape and len(a) > 0
        np.testing.assert_array_equal(a[np.lexsort(a.T)], b[np.lexsort(b.T)])

    # pipelined loop + shared host pool: same turns, same counts, same pool on both ranks
    assert r0["shared_pool"] and r1["shared_pool"] and not one["shared_pool"]
    for k, (a0, a1, a_one) in enumerate(zip(r0["pipelined"], r1["pipelined"], one["pipelined"])):
        assert a0[1:] == a1[1:] == a_one[1:]
        np.testing.assert_array_equal(a0[0], a1[0])
        # turn-major, and rank-major within a turn = draw order (the shards are contiguous): the
        # shared-pool path returns the very pool one rank returns, row for row
        np.testing.assert_array_equal(a0[0], a_one[0])
    for key in ("loop", "tail"):
        assert one[key][1:] == r0[key][1:] == r1[key][1:] and one[key][1] == 3 * 1001
        np.testing.assert_array_equal(r0[key][0], r1[key][0])  # the same pool on every rank
        same_rows(one[key][0], r0[key][0])  # rank-major instead of draw order, same rows
    # accumulate: the weights (hence the expected pool size each turn) do not depend on the
    # sharding; the uniforms of the rejection step do
    np.testing.assert_allclose(one["acc"][3], r0["acc"][3], rtol=1e-12)
    assert one["acc"][1] == r0["acc"][1] == r1["acc"][1]  # same number of turns
    np.testing.assert_array_equal(r0["acc"][0], r1["acc"][0])
    assert r0["acc"][2] == r1["acc"][2] >= 120 and len(r0["acc"][0]) == 120 == len(one["acc"][0])
    assert r0["acc"][4] == 512 and one["acc"][4] == 1024
    # the two ranks' counter blocks of a rejection step are adjacent and disjoint
    (b0, n0), (b1, n1) = r0["acc"][5][-1], r1["acc"][5][-1]
    assert n0 == n1 and b1 == b0 + n0


@pytest.mark.parametrize("accumulate", [False, True])
def test_nothing_survives_the_truncation(monkeypatch, accumulate):
    """flowproposal.py:436-439: turns whose rows are all truncated only count towards
    ``max_samples``; the loop ends there with an empty pool of the right dtype."""
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import GeneralPopulateEngine, PopulateEngine

    sim = _simdevice.install(monkeypatch)
    nf, D = _flow()
    names = [f"x{i}" for i in range(D)]
    scale, shift = _zscore(D)
    for cls in (PopulateEngine,) if accumulate else (PopulateEngine, GeneralPopulateEngine):
        eng = cls(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
        pre = (np.zeros(D, dtype=np.int32),) if cls is GeneralPopulateEngine else ()
        eng.configure(*pre, scale, shift, np.full(D, -4.0), np.full(D, 4.0), -1.0, 1e-3)  # |z| <= 0.001: nobody
        eng.seed = 3
        run = eng.run_accumulate if accumulate else eng.run
        rows, n_proposed, n_accepted = run(100, 500, max_samples=1600)
        assert (len(rows), n_proposed, n_accepted) == (0, 2000, 0) and rows.dtype == get_dtype(names)
    if accumulate:
        assert eng.last_accumulate["rejects"] == [] and eng.last_accumulate["n_expected"] == [0.0] * 4


def test_standalone_proposal_contract_on_simulated_device(tmp_path, monkeypatch):
    """``B200FlowProposal`` (the standalone mirror): the attributes the sampler relies on
     -- pool, indices, acceptance, draw()/repopulate, pool-size scaling,
    min_log_q preparation, pickling -- with the flow on the simulated device."""
    import pickle

    from nessai_b200.livepoint import numpy_array_to_live_points
    from nessai_b200.proposal import B200FlowProposal

    _simdevice.install(monkeypatch)
    nf, D = _flow()

    class SimModel(_simdevice.SimFlowModel):
        weights_file = None

        def __init__(self, flow_config=None, training_config=None, output=None, rng=None):
            super().__init__(nf, D)
            self.flow_config = flow_config

        def initialise(self):
            pass

        def forward_and_log_prob(self, x):
            z, lj = nf.forward(np.asarray(x, dtype=np.float64))
            return z, -0.5 * np.sum(z * z, axis=1) - 0.5 * D * np.log(2 * np.pi) + lj

    class Box:
        names = [f"x{i}" for i in range(D)]
        bounds = {n: [-4.0, 4.0] for n in names}

        def log_prior(self, x):
            a = np.stack([x[n] for n in self.names], axis=-1)
            return np.where(np.all((a >= -4) & (a <= 4), axis=-1), -D * np.log(8.0), -np.inf)

        def log_likelihood(self, x):
            return -0.5 * np.sum(np.stack([x[n] for n in self.names], axis=-1) ** 2, axis=-1)

    monkeypatch.setattr(B200FlowProposal, "_FlowModelClass", SimModel)
    prop = B200FlowProposal(Box(), rng=np.random.default_rng(3), flow_config=dict(n_blocks=4), output=str(tmp_path),
                            poolsize=300, drawsize=2000, truncation_methods=["latent_radius", "min_log_q"])
    with pytest.raises(RuntimeError):
        prop.populate(None, n_samples=10)  # flowproposal.py:401-405
    prop.initialise()
    live = numpy_array_to_live_points(1.5 * np.random.default_rng(5).standard_normal((500, D)) + 0.3, Box.names)
    prop.check_state(live)
    prop.training_data = live
    prop.populate(live[0], n_samples=300, max_samples=10**6)
    assert prop.populated and prop.samples.dtype == prop.x_dtype and prop.samples.size == 300
    assert sorted(prop.indices) == list(range(300)) and 0 < prop.population_acceptance <= 1
    assert prop._min_log_q == prop.forward_pass(live)[1].min()
    assert np.all(prop.forward_pass(prop.samples)[1] > prop._min_log_q - 1e-6)
    assert np.all(prop.samples["logP"] == -D * np.log(8.0)) and np.all(np.isfinite(prop.samples["logL"]))
    np.testing.assert_allclose(prop.samples["logL"], Box().log_likelihood(prop.samples))
    # draw() serves the pool in the permuted order and repopulates when it runs dry
    first = prop.samples[prop.indices.tolist()[-1]].copy()
    assert prop.draw(live[0]) == first and len(prop.indices) == 299
    for _ in range(299):
        prop.draw(live[0])
    assert not prop.populated and prop.populated_count == 1
    prop.ns_acceptance = 0.25
    prop.draw(live[0])  # pool-size scaling (flowproposal/base.py:416-435): 1 / acceptance
    assert prop.populated_count == 2 and prop.poolsize == 1200 and prop.samples.size == 1200
    # pickling drops the device objects (flowproposal/base.py:1286-1309)
    state = pickle.loads(pickle.dumps(prop.__getstate__()))
    assert "flow" not in state and "_engine" not in state and "model" not in state and state["resume_populated"]


def test_gpu_test_bodies_pass_on_the_simulated_device():
    """tests/test_gpu_zz_tail_accumulate.py was written without access to a GPU: its test bodies
    (index arithmetic, loop replay, tolerances) are executed here against the simulated device,
    in a subprocess because the dry run patches torch globally.  The bodies of the existing
    populate tests (tests/test_gpu_populate.py, GPU-verified earlier) run too: a regression guard
    for the shared host code the new paths hook into."""
    import json
    import os
    import subprocess
    import sys

    from conftest import REPO

    res = subprocess.run([sys.executable, os.path.join(REPO, "tests", "tools", "dryrun_gpu_tests_on_sim.py")],
                         capture_output=True, text=True, cwd=REPO, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert json.loads(res.stdout.strip().splitlines()[-1]) == {"dryrun_failed": 0, "of": 15}


def test_pipelined_loop_logic_equals_serial_loop(monkeypatch):
    """The software-pipelined ``PopulateEngine.run`` (speculative next draw, counts through a
    pinned buffer + event, records copied ahead of their count, pinned destination sized from a
    hint) against the one-synchronisation-per-turn loop over many sizes / hints / stop conditions.
    On the simulated device every "asynchronous" operation completes at once, so this checks the
    bookkeeping (what tests/test_gpu_populate.py::test_pipelined_loop_matches_serial_loop checks on
    the GPU for four cases), not the stream ordering."""
    from nessai_b200 import proposal
    from nessai_b200.livepoint import get_dtype

    pipelined = proposal.PopulateEngine.run  # before install() replaces it by the serial loop
    _simdevice.install(monkeypatch)
    _simdevice.install_fake_cuda_async(monkeypatch)
    nf, D = _flow()
    names = [f"x{i}" for i in range(D)]
    scale, shift = _zscore(D)

    def engine():
        e = proposal.PopulateEngine(_simdevice.SimFlowModel(nf, D), names, get_dtype(names))
        e.configure(scale, shift, np.full(D, -4.0), np.full(D, 4.0), -D * np.log(8.0), 4.9)
        e.seed = 99
        return e

    ea, eb = engine(), engine()
    rng = np.random.default_rng(0)
    drawsize = 600  # ~20 rows accepted per turn
    cases = [(50, 10**6, None), (1, 10**6, 1), (10**5, 4 * drawsize - 1, None), (45, 10**6, 10**4), (200, 10**6, 3)]
    cases += [(int(rng.integers(1, 300)), int(rng.choice([10**6, drawsize * int(rng.integers(1, 8))])),
               [None, 1, int(rng.integers(1, 400))][int(rng.integers(0, 3))]) for _ in range(25)]
    for n_samples, max_samples, hint in cases:
        if hint is not None:
            ea._accept_hint = hint
        ra, pa, aa = pipelined(ea, n_samples, drawsize, max_samples=max_samples)
        rb, pb, ab = eb._run_serial(n_samples, drawsize, max_samples, None, False)
        assert (pa, aa) == (pb, ab) and ea._turn_rows == eb._turn_rows, (n_samples, max_samples, hint)
        assert ra.dtype == rb.dtype and len(ra) == len(rb) == min(aa, n_samples)
        assert ra.tobytes() == rb.tobytes(), (n_samples, max_samples, hint)


# ------------------------------------------------------------------ engines + the product's CUDA sources, on the CPU
@pytest.fixture(scope="module")
def simt_libs(tmp_path_factory):
    import _simtdevice

    libs = _simtdevice.build(tmp_path_factory.mktemp("simtdevice"))
    if libs is None:
        pytest.skip("no g++ with C++20 <barrier>")
    return libs


def test_engines_over_the_cuda_sources_match_the_oracle_device(monkeypatch, simt_libs):
    """The real Python engines driving the product's own CUDA sources (generic draw kernel,
    non-affine tail, sum-exp, rejection + compaction: compiled for the CPU against the SIMT shim)
    return the pools the oracle-backed simulated device returns: the ordinary loop, the non-affine
    tail and accumulate_weights -- the whole populate path minus nvcc and the tcgen05
    specialisations, without a GPU."""
    import _simtdevice

    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import GeneralPopulateEngine, PopulateEngine
    from nessai_b200.spec import FlowSpec

    g, cfg, sd = load_golden("c2_realnvp_mlp")
    spec = FlowSpec(cfg)
    theta = np.zeros(spec.n_theta, np.float32)
    ints = {}
    spec.load_state_dict_numpy(sd, theta, ints)
    prog = spec.fold(theta, ints).program(True)
    nf, D = _flow()
    names = [f"x{i}" for i in range(D)]
    scale, shift = _zscore(D)
    lo, hi, lpc, radius, drawsize = np.full(D, -4.0), np.full(D, 4.0), -D * np.log(8.0), 4.9, 1500
    kind = (np.arange(D) % 4).astype(np.int32)
    sc = np.where(kind == 1, 8.0, np.where(kind == 3, 0.5, np.where(kind == 2, -2.0, 1.4)))
    sh = np.where(kind == 1, -4.0, np.where(kind == 2, 3.0, 0.1))
    lo2 = np.where(kind == 1, -4.0, np.where(kind == 2, -3.0, np.where(kind == 3, 0.0, -5.0)))
    hi2 = np.where(kind == 1, 4.0, np.where(kind == 2, 3.0, np.where(kind == 3, 6.0, 5.0)))

    def run_all(flow_model):
        out = {}
        eng = PopulateEngine(flow_model, names, get_dtype(names))
        eng.configure(scale, shift, lo, hi, lpc, radius, min_log_q=-40.0)
        eng.seed = 2024
        out["loop"] = eng.run(60, drawsize, max_samples=10**6)
        out["accumulate"] = eng.run_accumulate(80, drawsize, max_samples=10**6)
        gen = GeneralPopulateEngine(flow_model, names, get_dtype(names))
        gen.configure(kind, sc, sh, lo2, hi2, -3.0, radius, min_log_q=-27.0)
        gen.seed = 2025
        out["tail"] = gen.run(120, drawsize, max_samples=10**6)
        out["tail_accumulate"] = gen.run_accumulate(150, drawsize, max_samples=10**6)
        return out

    _simdevice.install(monkeypatch)
    ref = run_all(_simdevice.SimFlowModel(nf, D))  # the oracle-backed device
    sim = _simtdevice.install(monkeypatch, simt_libs)
    got = run_all(_simtdevice.SimtFlowModel(spec, prog))  # the CUDA sources
    assert {c[0] for c in sim.calls} == {"draw", "accept", "accept_x64", "tail", "sum_exp"}
    exact = 0
    for key in ref:
        (ra, pa, aa), (rb, pb, ab) = ref[key], got[key]
        exact += aa == ab
        # fp32 flow in the kernels, float64 in the oracle: a row at rounding distance from a
        # threshold may flip, everything else is the same pool
        assert pa == pb and abs(aa - ab) <= 2, key
        if aa == ab:
            assert len(ra) == len(rb) > 0
            a = np.stack([ra[nm] for nm in names], axis=-1)
            b = np.stack([rb[nm] for nm in names], axis=-1)
            np.testing.assert_allclose(b, a, rtol=2e-4, atol=2e-4, err_msg=key)
            np.testing.assert_array_equal(ra["logP"], rb["logP"])
    assert exact >= 3  # (flips are rare events; all four agree with these seeds)
