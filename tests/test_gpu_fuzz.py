"""Randomised parity sweep: flow configurations drawn over the reference's ``flow_config`` space
(/root/reference/src/nessai/flowmodel/config.py:12-24, flows/realnvp.py:76-214, flows/nsf.py:60-130,
flows/maf.py:62-104) -- features, conditioner width (the default 2 * n_inputs, narrower, wider than
the tensor-core kernels take), blocks, conditioner depth and type, linear transform, BatchNorm,
activation, coupling type -- each evaluated by whichever kernel the library picks (tcgen05 wide /
narrow instantiations, or the generic fp32 interpreter) against the float64 oracle on the same
weights: inverse, forward, log-prob, round trip.  Seeds are fixed: the sweep is deterministic."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def draw_config(seed):
    rng = np.random.default_rng(1000 + seed)
    ftype = rng.choice(["realnvp", "realnvp", "realnvp", "nsf", "maf"])
    D = int(rng.integers(2, 21)) if ftype != "nsf" else int(rng.integers(2, 34))
    cfg = dict(n_inputs=D, n_blocks=int(rng.integers(1, 6)), n_layers=int(rng.integers(1, 4)), ftype=str(ftype))
    width = rng.choice(["default", "narrow", "sixty-four", "wide"], p=[0.4, 0.25, 0.25, 0.1])
    if width == "narrow":
        cfg["n_neurons"] = int(rng.integers(4, 33))
    elif width == "sixty-four":
        cfg["n_neurons"] = 64
    elif width == "wide":
        cfg["n_neurons"] = int(rng.integers(65, 97))
    if ftype == "realnvp":
        cfg["net"] = str(rng.choice(["resnet", "resnet", "mlp"]))
        cfg["linear_transform"] = [None, "lu", "lu", "permutation"][int(rng.integers(0, 4))]
        cfg["batch_norm_between_layers"] = bool(rng.integers(0, 2))
        cfg["use_volume_preserving"] = bool(rng.random() < 0.15)
        cfg["activation"] = str(rng.choice(["relu", "relu", "relu", "tanh", "swish"]))
    elif ftype == "nsf":
        cfg["batch_norm_between_layers"] = bool(rng.random() < 0.3)
        if rng.random() < 0.2:
            cfg["num_bins"] = int(rng.integers(4, 13))
    else:
        cfg["batch_norm_between_layers"] = bool(rng.integers(0, 2))
    return cfg


@pytest.mark.parametrize("seed", range(32))
def test_random_flow_config_matches_oracle(seed, tmp_path):
    from nessai_b200.flowmodel import B200FlowModel
    from test_oracle import numpy_flow

    cfg = draw_config(seed)
    D = cfg["n_inputs"]
    torch.manual_seed(seed)
    fm = B200FlowModel(flow_config=dict(cfg), training_config=dict(device_tag="cuda:0"), output=str(tmp_path))
    fm.initialise()
    rng = np.random.default_rng(seed)
    sd = {}
    for k, v in fm.model.state_dict().items():
        a = v.cpu().numpy()
        if a.dtype.kind == "f" and not k.endswith(".mask"):
            a = a + (0.04 * rng.standard_normal(a.shape)).astype(np.float32)
            if "running_var" in k:
                a = np.abs(a) + 0.5
        sd[k] = a
    fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    fm.model.eval()
    ocfg = dict(cfg)
    ocfg.setdefault("n_neurons", fm.model.spec.H)
    nf = numpy_flow(ocfg, sd)
    n = 777
    z = rng.normal(size=(n, D))
    x, logj = fm.inverse(z)
    x64, logj64 = nf.inverse(z.astype(np.float32).astype(np.float64))
    # the spline flow's fp32 evaluation is the loosest (tests/test_gpu_c3_nsf.py); MAF's inverse is D sequential passes
    tol = dict(rtol=3e-4, atol=3e-4) if cfg["ftype"] in ("nsf", "maf") else dict(rtol=1e-4, atol=1e-4)
    ok = np.isfinite(logj64) & (np.abs(x64).max(axis=1) < 50.0)  # (rows a random flow throws far out are ill-conditioned)
    assert ok.mean() > 0.9, cfg
    np.testing.assert_allclose(x[ok], x64[ok], err_msg=str(cfg), **tol)
    np.testing.assert_allclose(logj[ok], logj64[ok], err_msg=str(cfg), **tol)
    xs = x64[ok].astype(np.float32).astype(np.float64)
    zz, logp = fm.forward_and_log_prob(xs)
    np.testing.assert_allclose(logp, nf.log_prob(xs), err_msg=str(cfg), **tol)
    np.testing.assert_allclose(zz, z[ok], rtol=1e-3, atol=1e-3, err_msg=str(cfg))
    np.testing.assert_allclose(fm.log_prob(xs), logp, rtol=0, atol=0)


@pytest.mark.parametrize("seed", range(16))
def test_random_flow_config_training_step_matches_oracle(seed, tmp_path):
    """The same sweep through the fused training kernels: train-mode loss (batch-statistics
    BatchNorm, uncached LU) and every parameter gradient against the float64 training oracle
    (oracle/train_numpy.py), on a ragged batch, every second one weighted."""
    from nessai_b200.flowmodel import B200FlowModel
    from oracle.train_numpy import TrainStepOracle

    cfg = draw_config(100 + seed)
    D = cfg["n_inputs"]
    torch.manual_seed(seed)
    fm = B200FlowModel(flow_config=dict(cfg), training_config=dict(device_tag="cuda:0"), output=str(tmp_path))
    fm.initialise()
    rng = np.random.default_rng(seed)
    sd = {}
    for k, v in fm.model.state_dict().items():
        a = v.cpu().numpy()
        if a.dtype.kind == "f" and not k.endswith(".mask"):
            a = a + (0.04 * rng.standard_normal(a.shape)).astype(np.float32)
            if "running_var" in k:
                a = np.abs(a) + 0.5
        sd[k] = a
    fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    spec = fm.model.spec
    n_rows = int(rng.integers(40, 700))
    x = (1.2 * rng.standard_normal((n_rows, D)) + 0.3).astype(np.float32)
    w = rng.uniform(0.2, 2.0, size=n_rows).astype(np.float32) if seed % 2 else None
    theta64 = fm.model.theta_numpy().astype(np.float64)
    loss64, grad64 = TrainStepOracle(spec, fm.model.ints).loss_and_grad(theta64, x.astype(np.float64), weights=w)
    loss, grad, info = fm._trainer().loss_and_grad(torch.from_numpy(x).cuda(), None if w is None else torch.from_numpy(w).cuda())
    torch.cuda.synchronize()
    assert abs(float(loss) - loss64) < 3e-5 * max(1.0, abs(loss64)), cfg
    grad = grad.cpu().numpy().astype(np.float64)
    assert np.isfinite(grad).all(), cfg
    err = float(np.linalg.norm(grad - grad64) / max(np.linalg.norm(grad64), 1e-30))
    assert err < (5e-4 if cfg["ftype"] == "nsf" else 1e-4), (cfg, err)


@pytest.mark.parametrize("seed", range(12))
def test_random_flow_config_fused_populate_turn_matches_oracle(seed, tmp_path):
    """One fused populate turn (Philox draw -> radius truncation -> inverse flow -> float64 rescale ->
    bounds -> log-weights -> statistics, flowproposal.py:431-469) for RealNVP configurations of the
    sweep -- wide, narrow and generic draw kernels alike -- against the oracle driven by the same latents."""
    from nessai_b200.flowmodel import B200FlowModel
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import PopulateEngine
    from oracle.philox_numpy import latent_normals
    from test_oracle import numpy_flow

    rng = np.random.default_rng(500 + seed)
    s = 200 + seed
    cfg = draw_config(s)
    while cfg["ftype"] != "realnvp":
        s += 100
        cfg = draw_config(s)
    if seed % 3 == 0:  # make sure the narrow MLP draw kernel is in the sweep
        cfg.update(net="mlp", n_layers=2, activation="relu", n_neurons=int(rng.integers(4, 33)),
                   n_inputs=int(rng.integers(2, 17)))
    elif seed % 3 == 1:  # ... and the 17 .. 32-feature ResidualNet draw kernel
        cfg.update(net="resnet", n_layers=int(rng.integers(1, 3)), activation="relu", n_inputs=int(rng.integers(17, 33)))
        cfg.pop("n_neurons", None)
        if seed % 2:
            cfg["n_neurons"] = int(rng.integers(8, 33))
    D = cfg["n_inputs"]
    torch.manual_seed(seed)
    fm = B200FlowModel(flow_config=dict(cfg), training_config=dict(device_tag="cuda:0"), output=str(tmp_path))
    fm.initialise()
    sd = {}
    for k, v in fm.model.state_dict().items():
        a = v.cpu().numpy()
        if a.dtype.kind == "f":
            a = a + (0.04 * rng.standard_normal(a.shape)).astype(np.float32)
            if "running_var" in k:
                a = np.abs(a) + 0.5
        sd[k] = a
    fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    fm.model.eval()
    names = [f"x{i}" for i in range(D)]
    eng = PopulateEngine(fm, names, get_dtype(names))
    scale, shift = rng.uniform(0.5, 2.0, D), rng.uniform(-0.5, 0.5, D)
    lo, hi = np.full(D, -6.0), np.full(D, 6.0)
    r_max = float(np.sqrt(D) + 1.5)
    n = 5000
    eng.configure(scale, shift, lo, hi, -D * np.log(12.0), r_max)
    eng._ensure(n, n, True)
    eng.draw_turn(n, want_z=True)
    z = eng.d_z[:n].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(z, latent_normals(eng.seed, np.arange(n), D), atol=2e-5, rtol=1e-5)
    ocfg = dict(cfg)
    ocfg.setdefault("n_neurons", fm.model.spec.H)
    xp, lq = numpy_flow(ocfg, sd).sample_and_log_prob(z)
    x_ref = xp * scale + shift
    lq = lq - np.sum(np.log(scale))
    rad = np.sqrt(np.sum(z**2, axis=1))
    valid = (rad <= r_max) & np.all((x_ref >= lo) & (x_ref <= hi), axis=1) & np.isfinite(lq)
    edge = (np.abs(rad - r_max) < 1e-4) | np.any(np.abs(np.abs(x_ref) - 6.0) < 2e-3, axis=1)
    logq, logw = eng.d_logq[:n].cpu().numpy(), eng.d_logw[:n].cpu().numpy()
    dev_valid = ~np.isnan(logw)
    assert np.array_equal(dev_valid[~edge], valid[~edge]), cfg
    both = dev_valid & valid
    assert both.sum() > 0.2 * n, cfg
    np.testing.assert_allclose(eng.physical_x(n).cpu().numpy()[both], x_ref[both], rtol=1e-4, atol=1e-4, err_msg=str(cfg))
    np.testing.assert_allclose(logq[both], lq[both], rtol=1e-4, atol=1e-4, err_msg=str(cfg))
    np.testing.assert_allclose(logw[both], -D * np.log(12.0) - lq[both], rtol=1e-4, atol=1e-4)
    stats = eng.d_stats.cpu().numpy()
    assert stats[1] == dev_valid.sum() and stats[0] == logw[dev_valid].max()
