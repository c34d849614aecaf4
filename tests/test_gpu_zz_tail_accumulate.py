"""GPU parity of the two populate variants added after the round's last GPU session -- written
and CPU-checked (oracle pins, host-compiled row function, loop control) without access to a
GPU, so this file sorts last:

* the non-affine populate tail (``nb200_reparam_tail`` + ``nb200_populate_accept_x64``,
  ``GeneralPopulateEngine``): named pre- / post-rescalings (logit, log, exp, Gaussian CDF and
  its inverse) and boundary inversion, reparameterisations/rescale.py:263-291,570-590,635-660;
* ``accumulate_weights`` (``nb200_sum_exp``, ``PopulateEngine.run_accumulate``),
  flowproposal.py:414-417,471-490,504-512;
* ``B200AugmentedFlowProposal`` (plugin point P2 of proposal/augmented.py).

The bodies of the first two groups also run on the CPU against the simulated device
(tests/tools/dryrun_gpu_tests_on_sim.py, tests/test_host_loops_simulated.py).
"""

import numpy as np
import pytest
import torch
from conftest import load_golden, reference_or_skip

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("n,min_log_q,pre", [(5000, None, True), (100_003, -14.0, True), (5000, None, False),
                                             (1, None, True)])
def test_reparam_tail_kernel_matches_oracle(n, min_log_q, pre):
    """Every kind of h (identity, sigmoid, abs, exp, log, normal CDF, normal quantile) and the pair
    kinds of Angle (angle, angle mod 2 pi, radius, auxiliary radius with its chi prior), with and
    without the pre-affine map and the source permutation, saturated / overflowing / out-of-domain
    arguments and rows already dropped: the inputs of the host-compiled check
    (tests/test_reparam_oracle.py), on the GPU."""
    from test_reparam_oracle import TAIL_CASE, tail_case_inputs

    from nessai_b200 import _lib
    from oracle.reparam_numpy import tail_rows

    lib = _lib.load()
    c = {k: v.copy() for k, v in TAIL_CASE.items()}
    if not pre:
        c["pre_scale"], c["pre_shift"], c["src"] = None, None, None
        c["kind"][[7, 9]] = 0
    d = len(c["kind"])
    xp, logq_flow = tail_case_inputs(n)
    x_ref, lq_ref, lw_ref, valid = tail_rows(xp, logq_flow, log_prior_const=-2.5, min_log_q=min_log_q, **c)
    d_xp, d_kind = _dev(xp), _dev(c["kind"])
    d_pre = [_dev(c[k]) if pre else None for k in ("src", "pre_scale", "pre_shift")]
    d_c = [_dev(c[k]) for k in ("scale", "shift", "lo", "hi")]
    d_logq, d_logw = _dev(logq_flow.copy()), torch.empty(n, dtype=torch.float64, device="cuda")
    d_x64 = torch.empty((n, d), dtype=torch.float64, device="cuda")
    d_stats = _dev(np.array([-np.inf, 0.0]))
    _lib.check(lib.nb200_reparam_tail(
        n, d, d_xp.data_ptr(), d_kind.data_ptr(), *(a.data_ptr() if a is not None else None for a in d_pre),
        *(a.data_ptr() for a in d_c), -2.5, float("nan") if min_log_q is None else min_log_q,
        d_logq.data_ptr(), d_logw.data_ptr(), d_x64.data_ptr(), d_stats.data_ptr(), _stream()), "nb200_reparam_tail")
    torch.cuda.synchronize()
    logq, logw, x64, stats = (t.cpu().numpy() for t in (d_logq, d_logw, d_x64, d_stats))
    # device libm vs numpy / scipy: a few ulp in exp / log1p / erfc / erfcinv, so a row within
    # rounding of a bound or of min_log_q may flip
    with np.errstate(all="ignore"):
        edge = np.any((np.abs(x_ref - c["lo"]) < 1e-9) | (np.abs(x_ref - c["hi"]) < 1e-9), axis=1)
        if min_log_q is not None:
            edge |= np.abs(logq_flow - min_log_q) < 1e-6
    np.testing.assert_array_equal((~np.isnan(logw))[~edge], valid[~edge])
    np.testing.assert_array_equal(np.isnan(logq), np.isnan(logw))
    both = valid & ~np.isnan(logw)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(x64, x_ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(logq[both], lq_ref[both], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(logw[both], lw_ref[both], rtol=1e-11, atol=1e-11)
    ok = ~np.isnan(logw)
    assert stats[1] == ok.sum()
    if ok.any():
        assert stats[0] == logw[ok].max()


def test_sum_exp_and_x64_accept_match_numpy():
    from nessai_b200 import _lib
    from nessai_b200.livepoint import get_dtype
    from oracle.philox_numpy import accept_uniform

    lib = _lib.load()
    n, d, seed, offset = 70_001, 7, 987654321, 12345
    rng = np.random.default_rng(1)
    logw = rng.normal(-3.0, 0.5, size=n)
    logw[rng.random(n) < 0.3] = np.nan
    x64 = rng.normal(size=(n, d))
    valid = ~np.isnan(logw)
    mx = logw[valid].max()
    d_logw, d_x64, d_max = _dev(logw), _dev(x64), _dev(np.array([mx, 0.0]))
    parts = torch.full((296,), float("nan"), dtype=torch.float64, device="cuda")
    _lib.check(lib.nb200_sum_exp(d_logw.data_ptr(), n, d_max.data_ptr(), parts.data_ptr(), 296, _stream()), "sum_exp")
    got = float(parts.sum().cpu())
    np.testing.assert_allclose(got, np.exp(logw[valid] - mx).sum(), rtol=1e-12)
    # bit-reproducible: no atomics
    parts2 = torch.empty_like(parts)
    _lib.check(lib.nb200_sum_exp(d_logw.data_ptr(), n, d_max.data_ptr(), parts2.data_ptr(), 296, _stream()), "sum_exp")
    assert torch.equal(parts, parts2)
    # rejection step + compaction from float64 rows
    names = [f"x{i}" for i in range(d)]
    dtype = get_dtype(names)
    from nessai_b200.livepoint import empty_structured_array

    tmpl = _dev(empty_structured_array(1, dtype=dtype).view(np.uint8).copy())
    offs = np.asarray([dtype.fields[nm][1] for nm in names] + [dtype.fields["logP"][1]], dtype=np.int32)
    rows = torch.zeros(n * dtype.itemsize, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    scratch = torch.empty(n // 1024 + 2, dtype=torch.int64, device="cuda")
    cap = 3000  # 6134 rows pass the rejection step (computed with the numpy Philox)
    _lib.check(lib.nb200_populate_accept_x64(
        n, d, d_x64.data_ptr(), d_logw.data_ptr(), None, d_max.data_ptr(), seed, offset, -1.25,
        tmpl.data_ptr(), dtype.itemsize, offs.ctypes.data, int(dtype.fields["logL"][1]), rows.data_ptr(), cap, 0,
        counts.data_ptr(), scratch.data_ptr(), _stream()), "accept_x64")
    c = counts.cpu().numpy()
    u = accept_uniform(seed, offset + np.arange(n))
    margin = (logw - mx) - np.log(u)
    acc = valid & (margin > 0)
    assert not (valid & (np.abs(margin) < 1e-9)).any()
    assert c[0] == acc.sum() == 6134 and c[1] == cap
    rec = rows[: cap * dtype.itemsize].cpu().numpy().view(dtype)
    got = np.stack([rec[nm] for nm in names], axis=-1)
    np.testing.assert_array_equal(got, x64[acc][:cap])  # copied bit for bit, draw order
    assert np.all(rec["logP"] == -1.25) and np.all(np.isnan(rec["logL"])) and np.all(rec["it"] == 0)


def _general_engine(tmp_path, name="c2_realnvp_mlp", seed=77):
    from nessai_b200.flowmodel import B200FlowModel
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import GeneralPopulateEngine

    g, cfg, sd = load_golden(name)
    D = cfg["n_inputs"]
    fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=str(tmp_path))
    fm.initialise()
    fm.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    fm.model.eval()
    names = [f"x{i}" for i in range(D)]
    eng = GeneralPopulateEngine(fm, names, get_dtype(names))
    eng.seed = seed
    kind = np.arange(D) % 4  # identity, sigmoid, abs, exp in turn
    scale = np.where(kind == 1, 8.0, np.where(kind == 3, 0.5, np.where(kind == 2, -2.0, 1.4)))
    shift = np.where(kind == 1, -4.0, np.where(kind == 2, 3.0, 0.1))
    lo = np.where(kind == 1, -4.0, np.where(kind == 2, -3.0, np.where(kind == 3, 0.0, -5.0)))
    hi = np.where(kind == 1, 4.0, np.where(kind == 2, 3.0, np.where(kind == 3, 6.0, 5.0)))
    return eng, cfg, sd, (kind.astype(np.int32), scale, shift, lo, hi)


def test_general_engine_turn_and_loop_match_oracle(tmp_path):
    """draw kernel (identity map) + tail kernel + x64 rejection step against the float64 flow
    oracle followed by the tail oracle, on the same Philox rows."""
    from oracle.flow_numpy import NumpyFlow
    from oracle.philox_numpy import accept_uniform, latent_normals
    from oracle.reparam_numpy import tail_rows

    n = 20000
    eng, cfg, sd, (kind, scale, shift, lo, hi) = _general_engine(tmp_path)
    D = cfg["n_inputs"]
    lpc, r_max = -3.0, 4.9
    eng.configure(kind, scale, shift, lo, hi, lpc, r_max, 1.0, min_log_q=-27.0)  # cuts ~7 % of the rows
    eng._ensure(n, n, True)
    eng.draw_turn(n, want_z=True)
    z = eng.d_z[:n].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(z, latent_normals(eng.seed, np.arange(n), D), atol=2e-5, rtol=1e-5)
    x = eng.physical_x(n).cpu().numpy()
    logq, logw = eng.d_logq[:n].cpu().numpy(), eng.d_logw[:n].cpu().numpy()
    stats = eng.d_stats.cpu().numpy()
    nf = NumpyFlow(sd, ftype="realnvp", net="mlp", hidden_features=cfg["n_neurons"])
    xp64, lq_flow = nf.sample_and_log_prob(z)
    keep = np.sqrt(np.sum(z**2, axis=1)) <= r_max
    x_ref, lq_ref, lw_ref, valid = tail_rows(xp64, np.where(keep, lq_flow, np.nan), kind=kind, scale=scale,
                                             shift=shift, lo=lo, hi=hi, log_prior_const=lpc, min_log_q=-27.0)
    # rows within fp32 rounding of the radius, of a bound or of min_log_q may flip
    edge = (np.abs(np.sqrt(np.sum(z**2, axis=1)) - r_max) < 1e-4) | np.any(
        (np.abs(x_ref - lo) < 2e-3) | (np.abs(x_ref - hi) < 2e-3), axis=1) | (np.abs(lq_ref + 27.0) < 1e-3)
    dev_valid = ~np.isnan(logw)
    assert np.array_equal(dev_valid[~edge], valid[~edge]) and 0.02 * n < valid.sum() < 0.98 * n
    both = dev_valid & valid
    np.testing.assert_allclose(x[both], x_ref[both], rtol=2e-4, atol=2e-4)  # x' is fp32
    np.testing.assert_allclose(logq[both], lq_ref[both], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(logw[both], lw_ref[both], rtol=1e-4, atol=2e-4)
    assert stats[1] == dev_valid.sum() and stats[0] == logw[dev_valid].max()
    # rejection step: float64 rows copied into the records in draw order
    counts = eng.accept_turn(n, 0).cpu().numpy()
    u = accept_uniform(eng.seed, np.arange(n))
    margin = (logw - stats[0]) - np.log(u)
    acc = dev_valid & (margin > 0)
    if not (dev_valid & (np.abs(margin) < 1e-9)).any():
        assert counts[0] == acc.sum()
        rows = eng._gather_rows(int(counts[1]), n)
        got = np.stack([rows[nm] for nm in eng.names], axis=-1)
        np.testing.assert_array_equal(got, x[acc])
        assert np.all(rows["logP"] == lpc)
    # the whole (pipelined) loop of the base class over the overridden turn
    eng._turn_rows = 0
    # (~6 % of the draws are accepted: about five turns, the later ones drawn speculatively)
    rows, n_proposed, n_accepted = eng.run(5000, n, max_samples=400 * n)
    assert len(rows) == 5000 and n_accepted >= 5000 and n_proposed % n == 0 and n_proposed >= 3 * n
    a = np.stack([rows[nm] for nm in eng.names], axis=-1)
    assert np.all((a >= lo) & (a <= hi))
    # a serial run from the same counter returns the same bytes
    eng._turn_rows = 0
    rows2, p2, a2 = eng._run_serial(5000, n, 400 * n, None, False)
    assert (p2, a2) == (n_proposed, n_accepted) and rows2.tobytes() == rows.tobytes()


def test_identity_kinds_reproduce_the_fused_affine_path(tmp_path):
    """With every map affine the general engine (draw + tail + x64 accept) must agree with
    the fused affine tail of the draw kernel: same dropped rows, same weights, same pool."""
    from nessai_b200.proposal import PopulateEngine

    n = 30000
    eng, cfg, sd, _ = _general_engine(tmp_path, seed=5)
    D = cfg["n_inputs"]
    scale, shift = np.full(D, 1.3), np.linspace(-0.5, 0.5, D)
    lo, hi = np.full(D, -4.0), np.full(D, 4.0)
    lpc = -D * np.log(8.0)
    eng.configure(np.zeros(D, dtype=np.int32), scale, shift, lo, hi, lpc, 4.9)
    ref = PopulateEngine(eng.flow, eng.names, eng.row_dtype)
    ref.seed = eng.seed
    ref.configure(scale, shift, lo, hi, lpc, 4.9)
    eng.draw_turn(n)
    ref.draw_turn(n)
    lw_a, lw_b = eng.d_logw[:n].cpu().numpy(), ref.d_logw[:n].cpu().numpy()
    np.testing.assert_array_equal(np.isnan(lw_a), np.isnan(lw_b))
    ok = ~np.isnan(lw_a)
    np.testing.assert_allclose(lw_a[ok], lw_b[ok], rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(eng.physical_x(n).cpu().numpy(), ref.physical_x(n).cpu().numpy(), rtol=1e-14, atol=1e-14)
    np.testing.assert_array_equal(eng.d_stats.cpu().numpy()[1], ref.d_stats.cpu().numpy()[1])


def _affine_proposal(tmp_path, pool, **kw):
    from nessai_b200.livepoint import numpy_array_to_live_points
    from nessai_b200.proposal import B200FlowProposal

    g, cfg, sd = load_golden("c2_realnvp_mlp")
    d = cfg["n_inputs"]

    class Box:
        names = [f"x{i}" for i in range(d)]
        bounds = {n: [-4.0, 4.0] for n in names}

        def log_prior(self, x):
            a = np.stack([x[n] for n in self.names], axis=-1)
            return np.where(np.all((a >= -4) & (a <= 4), axis=-1), -d * np.log(8.0), -np.inf)

        def log_likelihood(self, x):
            return -0.5 * np.sum(np.stack([x[n] for n in self.names], axis=-1) ** 2, axis=-1)

    model = Box()
    torch.manual_seed(11)
    prop = B200FlowProposal(model, rng=np.random.default_rng(3), flow_config=cfg,
                            training_config=dict(device_tag="cuda:0"), output=str(tmp_path),
                            poolsize=pool, drawsize=pool, **kw)
    prop.initialise()
    live = numpy_array_to_live_points(1.5 * np.random.default_rng(5).standard_normal((500, d)) + 0.3, model.names)
    prop.check_state(live)
    prop.flow.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    prop.flow.model.eval()
    return prop, model, cfg, sd, live


@pytest.mark.parametrize("n_samples,max_turns", [(1200, 400), (10**6, 2)])
def test_accumulate_device_loop_matches_oracle(tmp_path, n_samples, max_turns):
    """run_accumulate: every turn lands in its slot with the draws of its own Philox rows, the
    expected pool size is the sum over all slots, the loop takes the reference's decisions and
    the final rejection step keeps samples[accept][:n_samples] in draw order."""
    from nessai_b200.proposal import AccumulateControl
    from oracle.flow_numpy import NumpyFlow
    from oracle.philox_numpy import accept_uniform, latent_normals
    from oracle.populate_numpy import populate_turn

    drawsize = 10_000
    max_samples = max_turns * drawsize - 1  # the loop ends after max_turns turns at the latest
    prop, model, cfg, sd, live = _affine_proposal(tmp_path, drawsize, accumulate_weights=True)
    D = cfg["n_inputs"]
    eng = prop._get_engine()
    eng.seed = 4242
    t0 = eng._turn_rows
    rows, n_proposed, n_accepted = eng.run_accumulate(n_samples, drawsize, max_samples=max_samples)
    info = eng.last_accumulate
    turns, stride = len(info["draw_offsets"]), info["stride"]
    assert stride == 10_016 and info["n_local"] == drawsize and n_proposed == turns * drawsize
    assert info["draw_offsets"][0] == t0 and 1 <= turns <= max_turns
    lw = eng.d_logw[: turns * stride].cpu().numpy()
    xs = (eng.d_xp[: turns * stride].to(torch.float64) * eng.d_scale + eng.d_shift).cpu().numpy()
    stats = eng.d_stats.cpu().numpy()
    valid = ~np.isnan(lw)
    assert stats[1] == valid.sum() and stats[0] == lw[valid].max()
    nf = NumpyFlow(sd, ftype="realnvp", net="mlp", hidden_features=cfg["n_neurons"])
    ctl = AccumulateControl(n_samples, max_samples)
    n_rej = 0
    for t in range(turns):
        slot = slice(t * stride, t * stride + drawsize)
        assert np.all(np.isnan(lw[t * stride + drawsize : (t + 1) * stride]))  # slot padding never counts
        z = latent_normals(eng.seed, info["draw_offsets"][t] + np.arange(drawsize), D)
        o = populate_turn(nf, z, scale=prop.scale, shift=prop.shift, lo=-4.0, hi=4.0,
                          log_prior_const=-D * np.log(8.0), r_max=prop.radius)
        edge = (np.abs(np.sqrt(np.sum(z**2, axis=1)) - prop.radius) < 1e-4) | np.any(
            np.abs(np.abs(o["x"]) - 4.0) < 1e-3, axis=1)
        assert np.array_equal(valid[slot][~edge], o["valid"][~edge])
        both = valid[slot] & o["valid"]
        np.testing.assert_allclose(lw[slot][both], o["log_w"][both], rtol=1e-4, atol=1e-4)
        # the loop control, replayed on the device's own weights
        upto = lw[: (t + 1) * stride]
        v = ~np.isnan(upto)
        mx = upto[v].max()
        n_expected = np.exp(upto[v] - mx).sum()
        np.testing.assert_allclose(info["n_expected"][t], n_expected, rtol=1e-10)
        assert ctl.go_on()
        if ctl.turn_drawn(drawsize, bool(valid[slot].any()), info["n_expected"][t]):
            base, nrows = info["rejects"][n_rej]
            n_rej += 1
            assert nrows == (t + 1) * stride
            ctl.rejected(int((v & ((upto - mx) > np.log(accept_uniform(eng.seed, base + np.arange(nrows))))).sum()))
        ctl.end_turn()
    assert not ctl.go_on()
    if ctl.stale:
        n_rej += 1
    assert n_rej == len(info["rejects"])
    # the last rejection step is the pool
    base, nrows = info["rejects"][-1]
    assert nrows == turns * stride
    margin = (lw - stats[0]) - np.log(accept_uniform(eng.seed, base + np.arange(nrows)))
    acc = valid & (margin > 0)
    if not (valid & (np.abs(margin) < 1e-9)).any():
        assert n_accepted == acc.sum() and len(rows) == min(n_accepted, n_samples)
        got = np.stack([rows[nm] for nm in model.names], axis=-1)
        np.testing.assert_allclose(got, xs[acc][:n_samples], rtol=1e-12, atol=1e-14)
    if n_samples == 1200:
        assert n_accepted >= n_samples and turns >= 2  # acceptance of this flow is ~3 %: several turns
    else:
        assert turns == max_turns and len(rows) < n_samples  # ended by max_samples
    # fresh uniforms for every rejection step: the counter blocks do not overlap
    spans = sorted((b, b + r) for b, r in info["rejects"])
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))


def test_accumulate_through_the_standalone_proposal(tmp_path):
    drawsize = 20_000
    prop, model, cfg, sd, live = _affine_proposal(tmp_path, drawsize, accumulate_weights=True)
    prop.populate(live[0], n_samples=2000, max_samples=200 * drawsize)
    assert prop.populated and len(prop.samples) == 2000 and len(prop.indices) == 2000
    a = np.stack([prop.samples[n] for n in model.names], axis=-1)
    assert np.all((a >= -4) & (a <= 4)) and np.all(np.isfinite(prop.samples["logL"]))
    assert 0 < prop.population_acceptance < 1
    # the pool follows the weights: the importance-weighted mean of the draws predicts the pool mean
    eng = prop._engine
    turns, stride = len(eng.last_accumulate["draw_offsets"]), eng.last_accumulate["stride"]
    lw = eng.d_logw[: turns * stride].cpu().numpy()
    xs = (eng.d_xp[: turns * stride].to(torch.float64) * eng.d_scale + eng.d_shift).cpu().numpy()
    ok = ~np.isnan(lw)
    w = np.exp(lw[ok] - lw[ok].max())
    mean_w = (xs[ok] * w[:, None]).sum(0) / w.sum()
    ess = w.sum() ** 2 / (w**2).sum()
    sd_w = np.sqrt(((xs[ok] - mean_w) ** 2 * w[:, None]).sum(0) / w.sum())
    assert np.all(np.abs(a.mean(0) - mean_w) < 6 * sd_w * np.sqrt(1 / 2000 + 1 / ess))


def test_general_accumulate_reproduces_affine_accumulate(tmp_path):
    """accumulate_weights over the non-affine tail (slot-offset tail launches, scratch statistics
    for the draw kernel, float64-row rejection step): with identity maps it must be the affine
    engine's accumulating loop -- same turns, same expected pool sizes, same pool."""
    from nessai_b200.proposal import PopulateEngine

    drawsize = 20_000
    gen, cfg, sd, _ = _general_engine(tmp_path, seed=31)
    D = cfg["n_inputs"]
    scale, shift = np.full(D, 1.3), np.linspace(-0.5, 0.5, D)
    lo, hi, lpc = np.full(D, -4.0), np.full(D, 4.0), -D * np.log(8.0)
    gen.configure(np.zeros(D, dtype=np.int32), scale, shift, lo, hi, lpc, 4.9, min_log_q=-40.0)
    aff = PopulateEngine(gen.flow, gen.names, gen.row_dtype)
    aff.seed = gen.seed
    aff.configure(scale, shift, lo, hi, lpc, 4.9, min_log_q=-40.0)
    ra, pa, aa = aff.run_accumulate(3000, drawsize, max_samples=10**7)
    rg, pg, ag = gen.run_accumulate(3000, drawsize, max_samples=10**7)
    assert pa == pg and aff.last_accumulate["rejects"] == gen.last_accumulate["rejects"]
    np.testing.assert_allclose(aff.last_accumulate["n_expected"], gen.last_accumulate["n_expected"], rtol=1e-9)
    assert abs(aa - ag) <= 2 and len(ra) == len(rg) == 3000  # (a weight within an ulp of log u may flip)
    if aa == ag:
        for nm in gen.names:
            np.testing.assert_allclose(rg[nm], ra[nm], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["c2_realnvp_mlp", "c2_realnvp_resnet", "d6_nsf", "d8_maf", "c1_realnvp_2d"])
def test_reference_self_consistency_pins(name, tmp_path):
    """The self-consistency properties the reference's own tests pin at this boundary
    (SURVEY.md 8c), on the kernels:
    forward_and_log_prob(x) == (forward(x)[0], log_prob(x)) bit for bit
    (tests/test_flows/test_included_flows.py:114-126); float64 outputs
    (tests/test_flowmodel/test_flowmodel_base.py:543-549); sample_and_log_prob(z=z) ==
    base_log_prob(z) - inverse(z)[1] (:552-570); outputs identical before and after the
    train() -> eval() cache reset (:751-788, bit-exact)."""
    from test_gpu_flow import make_model

    g, cfg, sd = load_golden(name)
    fm = make_model(cfg, sd, tmp_path)
    x, zz = np.asarray(g["x"], dtype=np.float64), np.asarray(g["z"], dtype=np.float64)
    z, lp = fm.forward_and_log_prob(x)
    assert z.dtype == np.float64 and lp.dtype == np.float64 and z.shape == x.shape and lp.shape == (len(x),)
    xt = fm.numpy_array_to_tensor(x)
    z2, _ = fm.model.forward(xt)
    lp2 = fm.model.log_prob(xt)
    np.testing.assert_array_equal(z, z2.cpu().numpy().astype(np.float64))
    np.testing.assert_array_equal(lp, lp2.cpu().numpy().astype(np.float64))
    np.testing.assert_array_equal(fm.log_prob(x), lp)
    x3, lq = fm.sample_and_log_prob(z=zz)
    x4, lj = fm.inverse(zz)
    assert x3.dtype == lq.dtype == x4.dtype == lj.dtype == np.float64
    base = fm.model.base_distribution_log_prob(fm.numpy_array_to_tensor(zz)).cpu().numpy().astype(np.float64)
    np.testing.assert_array_equal(x3, x4)
    ok = np.isfinite(lq)
    assert ok.mean() > 0.9
    np.testing.assert_allclose(lq[ok], (base - lj)[ok], rtol=1e-5, atol=1e-4)
    fm.model.train()
    fm.model.eval()  # flowmodel/base.py:680-682
    z5, lp5 = fm.forward_and_log_prob(x)
    np.testing.assert_array_equal(z5, z)
    np.testing.assert_array_equal(lp5, lp)
    # inputs are caller-owned and never mutated; outputs are fresh and writable
    assert np.array_equal(x, np.asarray(g["x"], dtype=np.float64))
    lp -= 1.0
    assert not np.shares_memory(lp, z)


@pytest.mark.reference
@pytest.mark.parametrize("variant", ["logit_and_default", "accumulate", "periodic_angle", "to_cartesian",
                                     "unit_hypercube"])
def test_flowsampler_variants_run_on_the_device_loop(tmp_path, variant):
    """The reference's FlowSampler, unmodified, with a logit + rescale-to-bounds
    reparameterisation (GeneralPopulateEngine), with accumulate_weights=True (run_accumulate)
    and with an Angle / ToCartesian reparameterisation (Cartesian pair + auxiliary radius with its
    chi prior, GeneralPopulateEngine over three flow features): the device loop is the one that runs."""
    reference_or_skip()
    from nessai.flowsampler import FlowSampler
    from test_gpu_nessai_plugin import make_model

    from nessai_b200.nessai_plugin import B200NessaiFlowProposal
    from nessai_b200.proposal import GeneralPopulateEngine, PopulateEngine

    kw = dict(logit_and_default=dict(reparameterisations={"x": "logit", "y": "default"}),
              accumulate=dict(accumulate_weights=True),
              periodic_angle=dict(reparameterisations={"x": "periodic", "y": "default"}),
              to_cartesian=dict(reparameterisations={"x": "to-cartesian", "y": "default"}),
              unit_hypercube=dict(map_to_unit_hypercube=True))[variant]
    model = make_model()
    if variant == "unit_hypercube":  # the two maps the reference asks the user for (model.py:603-614)
        class WithHypercube(type(model)):
            def to_unit_hypercube(self, x):
                u = x.copy()
                for n in self.names:
                    u[n] = (x[n] - self.bounds[n][0]) / (self.bounds[n][1] - self.bounds[n][0])
                return u

            def from_unit_hypercube(self, u):
                x = u.copy()
                for n in self.names:
                    x[n] = u[n] * (self.bounds[n][1] - self.bounds[n][0]) + self.bounds[n][0]
                return x

        model = WithHypercube()
    fs = FlowSampler(
        model, output=str(tmp_path), resume=False, seed=1234, nlive=200, plot=False,
        flow_proposal_class=B200NessaiFlowProposal, flow_config=dict(n_blocks=2),
        training_config=dict(max_epochs=50, patience=10), maximum_uninformed=200,
        max_iteration=700, poolsize=2000, checkpointing=False, **kw,
    )
    fs.run(plot=False, save=False)
    prop = fs.ns._flow_proposal
    assert prop.training_count >= 1 and prop.populated_count >= 1
    if variant == "logit_and_default":
        assert type(prop._engine) is GeneralPopulateEngine
    elif variant in ("periodic_angle", "to_cartesian"):
        assert type(prop._engine) is GeneralPopulateEngine and prop._engine.names == ["x", "y", "x_radial"]
        assert prop.flow.model.spec.D == 3 and prop.samples.dtype.names[:2] == ("x", "y")
        assert "x_radial" in prop.x.dtype.names and "x_radial" not in prop.samples.dtype.names
    elif variant == "unit_hypercube":
        assert type(prop._engine) is PopulateEngine and prop.map_to_unit_hypercube
        u = np.stack([prop.x[n] for n in ("x", "y")], axis=-1)
        assert np.all((u >= 0) & (u < 1)) and np.all(np.abs(prop.samples["x"]) <= 10)
    else:
        assert type(prop._engine) is PopulateEngine and getattr(prop._engine, "last_accumulate", None)
    # truncated at max_iteration (the reference's own CPU proposal gives -7.3 here; analytic -5.99)
    assert np.isfinite(fs.ns.log_evidence) and -9.5 < fs.ns.log_evidence < -4.0


@pytest.mark.reference
@pytest.mark.parametrize("marginalise", [False, True])
def test_augmented_flow_proposal_on_b200_flows(tmp_path, marginalise):
    """``AugmentedFlowProposal`` (proposal/augmented.py) with plugin point P2 swapped: the flow
    over dims + augment_dims inputs with the proposal's custom mask is trained and evaluated by
    the kernels, the marginalisation as one batch of n * n_marg rows."""
    reference_or_skip()
    from nessai.flowsampler import FlowSampler
    from test_gpu_nessai_plugin import make_model

    from nessai_b200.flowmodel import B200FlowModel
    from nessai_b200.nessai_plugin import B200AugmentedFlowProposal

    fs = FlowSampler(
        make_model(), output=str(tmp_path), resume=False, seed=1234, nlive=200, plot=False,
        flow_proposal_class=B200AugmentedFlowProposal, flow_config=dict(n_blocks=2),
        training_config=dict(max_epochs=50, patience=10), maximum_uninformed=200,
        max_iteration=500, poolsize=1000, checkpointing=False,
        augment_dims=1, marginalise_augment=marginalise, n_marg=8,
    )
    fs.run(plot=False, save=False)
    prop = fs.ns._flow_proposal
    assert isinstance(prop, B200AugmentedFlowProposal) and isinstance(prop.flow, B200FlowModel)
    assert prop.flow.model.spec.D == 3 and list(prop.flow_config["mask"]) == [1.0, 1.0, -1.0]
    assert prop.training_count >= 1 and prop.populated_count >= 1
    assert np.isfinite(fs.ns.log_evidence)
    if not marginalise:
        # the populate loop ran on the device: the augment parameter is a field of the population records
        # (kind 17 of the tail: identity + its N(0, 1) prior) that never reaches the sampler
        from nessai_b200.proposal import GeneralPopulateEngine

        assert type(prop._engine) is GeneralPopulateEngine and prop._engine.names == ["x", "y", "e_0"]
        assert "e_0" in prop.x.dtype.names and "e_0" not in prop.samples.dtype.names
    else:
        assert prop._engine is None  # the marginal estimate of log q: the reference's loop over our flow kernels
    # the marginalised density of a point agrees with a brute-force estimate from the same flow
    if marginalise:
        x = np.array([[0.3, -0.2, 0.0]] * 4)
        prop.n_marg = 4000
        prop.rng = np.random.default_rng(0)
        a = prop._marginalise_augment(x.copy())
        prop.rng = np.random.default_rng(1)
        b = prop._marginalise_augment(x.copy())
        assert np.all(np.isfinite(a)) and np.max(np.abs(a - b)) < 0.3
