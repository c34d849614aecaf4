"""Parity of the CUDA flow (through the C ABI, via B200FlowModel) with the
reference's outputs (golden fixtures) and the float64 oracle.

Stated tolerance (BASELINE.json north_star): log_prob / log|J| rtol 1e-4 in fp32;
samples x/z: atol 1e-4 + rtol 1e-4.
"""

import pickle

import numpy as np
import pytest
import torch
from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-4


def make_model(cfg, sd, tmp_path):
    from nessai_b200.flowmodel import B200FlowModel

    fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=str(tmp_path))
    fm.initialise()
    fm.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    fm.model.eval()
    return fm


def test_inverse_and_forward_match_reference(golden, tmp_path):
    name, g, cfg, sd = golden
    fm = make_model(cfg, sd, tmp_path)
    x, logj = fm.inverse(g["z"])
    assert x.dtype == np.float64 and logj.dtype == np.float64 and x.flags.writeable
    np.testing.assert_allclose(x, g["inv_x"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(logj, g["inv_logj"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(x, g["inv_x64"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(logj, g["inv_logj64"], rtol=RTOL, atol=ATOL)
    x2, logq = fm.sample_and_log_prob(z=g["z"])
    np.testing.assert_allclose(x2, x, rtol=0, atol=0)
    np.testing.assert_allclose(logq, g["inv_logq"], rtol=RTOL, atol=ATOL)
    z, logp = fm.forward_and_log_prob(g["x"])
    np.testing.assert_allclose(z, g["fwd_z"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(logp, g["fwd_logprob"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(logp, g["fwd_logprob64"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(fm.log_prob(g["x"]), logp, rtol=0, atol=0)
    zt, lj = fm.model.forward(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(lj.cpu().numpy(), g["fwd_logj"], rtol=RTOL, atol=ATOL)


def test_round_trip_property(golden, tmp_path):
    """inverse(forward(x)) == x, logJ_fwd == -logJ_inv to 5 decimals
    (/root/reference/tests/test_flows/test_included_flows.py:144-154)."""
    name, g, cfg, sd = golden
    fm = make_model(cfg, sd, tmp_path)
    zt, lj = fm.model.forward(torch.from_numpy(g["x"]))
    xt, ilj = fm.model.inverse(zt)
    np.testing.assert_array_almost_equal(xt.cpu().numpy(), g["x"], decimal=4)
    np.testing.assert_array_almost_equal(lj.cpu().numpy(), -ilj.cpu().numpy(), decimal=4)


@pytest.mark.parametrize("n", [0, 1, 31, 127, 129, 1000])
def test_ragged_sizes(n, tmp_path):
    g, cfg, sd = load_golden("c2_realnvp_mlp")
    fm = make_model(cfg, sd, tmp_path)
    z = np.resize(g["z"], (n, 16)) if n else np.zeros((0, 16))
    x, lq = fm.sample_and_log_prob(z=z)
    assert x.shape == (n, 16) and lq.shape == (n,)
    if n:
        k = min(n, len(g["z"]))
        np.testing.assert_allclose(lq[:k], g["inv_logq"][:k], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", ["c2_realnvp_mlp", "c2_realnvp_resnet", "c1_realnvp_2d", "d6_nsf", "ac_20d"])
def test_large_batch_is_consistent_with_small(name, tmp_path):
    """Size-independent property at the bench size: 1e6 rows = tiled copies of a few hundred rows must
    give exactly the small batch's result, row for row -- whatever tile, CTA, layer pass or scratch
    hand-off a row goes through (MLP kernel: one pass; ResidualNet: two passes through the scratch
    buffer, or one in the narrow instantiation; spline and 17 .. 32-feature kernels: one launch per
    layer), in both directions."""
    if name == "ac_20d":  # RealNVP, default conditioner, 20 features: the affine-coupling tile kernel
        from nessai_b200.flowmodel import B200FlowModel

        torch.manual_seed(5)
        fm = B200FlowModel(flow_config=dict(n_inputs=20, n_blocks=3, ftype="realnvp"),
                           training_config=dict(device_tag="cuda:0"), output=str(tmp_path))
        fm.initialise()
        # (nflows' BatchNorm starts with running_var = 0: an untrained flow in eval mode scales by 316 per layer)
        sd0 = {k: (torch.ones_like(v) if k.endswith("running_var") else v) for k, v in fm.model.state_dict().items()}
        fm.model.load_state_dict(sd0)
        fm.model.eval()
        zs = np.random.default_rng(5).normal(size=(300, 20))
    else:
        g, cfg, sd = load_golden(name)
        fm = make_model(cfg, sd, tmp_path)
        zs = np.asarray(g["z"], dtype=np.float64)
    n = 1_000_000
    idx = np.arange(n) % len(zs)
    x0, lq0 = fm.sample_and_log_prob(z=zs)
    x, lq = fm.sample_and_log_prob(z=zs[idx])
    np.testing.assert_array_equal(lq, lq0[idx])
    np.testing.assert_array_equal(x, x0[idx])
    ok = np.isfinite(lq0)
    z0, lp0 = fm.forward_and_log_prob(x0[ok])
    k = np.arange(n) % int(ok.sum())
    z1, lp1 = fm.forward_and_log_prob(x0[ok][k])
    np.testing.assert_array_equal(lp1, lp0[k])
    np.testing.assert_array_equal(z1, z0[k])
    np.testing.assert_allclose(z0, zs[ok], rtol=2e-3, atol=2e-3)  # ... and the round trip closes


def test_latent_sampling_matches_philox_oracle(tmp_path):
    from oracle.philox_numpy import latent_normals

    g, cfg, sd = load_golden("c2_realnvp_mlp")
    fm = make_model(cfg, sd, tmp_path)
    z = fm.sample_latent_distribution(5000)
    assert z.shape == (5000, 16) and z.dtype == np.float64
    seed = fm.model._rng_seed
    ref = latent_normals(seed, np.arange(5000), 16)
    np.testing.assert_allclose(z, ref, atol=2e-5, rtol=1e-5)
    z2 = fm.sample_latent_distribution(5000)  # next rows of the stream
    ref2 = latent_normals(seed, np.arange(5000, 10000), 16)
    np.testing.assert_allclose(z2, ref2, atol=2e-5, rtol=1e-5)
    big = fm.sample_latent_distribution(400_000)
    assert abs(big.mean()) < 5e-3 and abs(big.var() - 1) < 5e-3


def test_sample_and_log_prob_self_consistent(tmp_path):
    """test_included_flows.py:129-141."""
    g, cfg, sd = load_golden("c2_realnvp_resnet")
    fm = make_model(cfg, sd, tmp_path)
    x, lq = fm.sample_and_log_prob(N=2000)
    lp = fm.log_prob(x)
    ok = np.isfinite(lq)
    np.testing.assert_allclose(lp[ok], lq[ok], rtol=1e-3, atol=1e-3)


def test_weights_roundtrip_and_pickle(tmp_path):
    g, cfg, sd = load_golden("c1_realnvp_2d")
    fm = make_model(cfg, sd, tmp_path)
    f = str(tmp_path / "model.pt")
    fm.save_weights(f)
    loaded = torch.load(f, weights_only=True)
    assert list(loaded) == list(sd)
    for k in sd:
        np.testing.assert_array_equal(loaded[k].numpy(), sd[k])
    fm2 = make_model(cfg, {k: np.zeros_like(v) if v.dtype.kind == "f" else v for k, v in sd.items()}, tmp_path)
    fm2.load_weights(f)
    np.testing.assert_array_equal(fm2.inverse(g["z"])[0], fm.inverse(g["z"])[0])
    state = pickle.loads(pickle.dumps(fm))
    assert state.initialised is False and "model" not in state.__dict__ and "_optimiser" not in state.__dict__


@pytest.mark.parametrize("name", ["c2_realnvp_mlp", "c2_realnvp_resnet", "d6_nsf", "d8_maf"])
def test_training_tracks_reference_history(name, tmp_path):
    """Same seeds, same data, same hyper-parameters as the golden run of the
    reference FlowModel.train (RealNVP with both conditioners, the spline flow and the
    masked autoregressive flow): bit-identical initial weights, and the first epochs'
    losses must agree."""
    from nessai_b200.flowmodel import B200FlowModel

    g, cfg, sd = load_golden(name)
    seed = 20251017
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    fm = B200FlowModel(flow_config=cfg, training_config=dict(max_epochs=8, patience=8, batch_size=1000),
                       output=str(tmp_path), rng=rng)
    fm.initialise()
    init = fm.model.state_dict()
    for k in init:
        np.testing.assert_array_equal(init[k].numpy(), g[f"init/{k}"])
    hist = fm.train(g["train_data"], plot=False)
    assert set(hist) == {"loss", "val_loss"} and len(hist["loss"]) == 8
    np.testing.assert_allclose(hist["loss"][:4], g["loss"][:4], rtol=2e-3)
    np.testing.assert_allclose(hist["val_loss"][:4], g["val_loss"][:4], rtol=5e-3)
    assert hist["loss"][-1] < hist["loss"][0]
    # trained flow evaluates through the kernels and stays self-consistent
    x, lq = fm.sample_and_log_prob(N=1000)
    lp = fm.log_prob(x)
    ok = np.isfinite(lq)
    np.testing.assert_allclose(lp[ok], lq[ok], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("name,launches", [("c2_realnvp_mlp", 1), ("c2_realnvp_resnet", 2)])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 3000, 100_003])
def test_tensor_core_paths_match_generic_kernel(name, launches, n, tmp_path):
    """The tcgen05 specialisations (MLP: one launch; ResidualNet: one launch per
    layer pass) against the generic fp32 interpreter on the same rows, both
    directions, ragged tile counts -- and the launch count proves which path ran."""
    from nessai_b200 import _lib

    g, cfg, sd = load_golden(name)
    fm = make_model(cfg, sd, tmp_path)
    lib = _lib.load()
    rng = np.random.default_rng(n)
    z = torch.from_numpy(rng.normal(size=(n, 16)).astype(np.float32)).cuda()
    out = {}
    try:
        for tc in (1, 0):
            lib.nb200_set_tensor_core_path(tc)
            _lib.reset_launch_count()
            x, logj, logq = fm.model._inverse(z)
            n_launch = _lib.launch_count()
            zz, flogj, logp = fm.model._forward(x)
            torch.cuda.synchronize()
            out[tc] = [t.cpu().numpy() for t in (x, logj, logq, zz, flogj, logp)]
            assert n_launch == (launches if tc else 1)
    finally:
        lib.nb200_set_tensor_core_path(1)
    for a, b in zip(out[1], out[0]):
        assert np.isfinite(a).all()
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL)
    # round trip through the tensor-core kernels
    np.testing.assert_allclose(out[1][3], z.cpu().numpy(), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("ftype,net,D,H,n_layers,launches,act", [(*c, "relu") if len(c) == 6 else c for c in [
    ("realnvp", "mlp", 16, 32, 2, 1),     # nessai's default width: 2 * n_inputs
    ("realnvp", "mlp", 4, 8, 2, 1),
    ("realnvp", "resnet", 16, 32, 2, 1),  # the default conditioner at its default width: 27 KB of weights a
    ("realnvp", "resnet", 7, 14, 2, 1),   # layer in the narrow image, the whole flow in one pass
    ("realnvp", "resnet", 12, 40, 1, 1),
    ("realnvp", "resnet", 8, 16, 3, 1),   # three residual blocks (wide image: 113 KB of weights a layer, one layer a pass)
    ("realnvp", "resnet", 16, 64, 3, 3),
    ("nsf", "resnet", 10, 20, 2, 3),
    ("nsf", "resnet", 32, 48, 2, 3),
    # RealNVP with the default conditioner at 17 .. 32 features: the tile machinery of the spline kernel
    # with an affine coupling (one launch per layer)
    ("realnvp", "resnet", 20, 40, 2, 3),
    ("realnvp", "resnet", 32, 64, 2, 3),
    ("realnvp", "resnet", 17, 34, 1, 3),
    ("realnvp", "resnet", 24, 30, 2, 3),
    ("realnvp", "resnet", 19, 38, 3, 3),
    # tanh / SiLU conditioners (flows/utils.py:200-205): the activation is a compile-time parameter of the
    # MLP and ResidualNet kernels (two SFU operations per value in front of the fp16 split)
    ("realnvp", "mlp", 16, 64, 2, 1, "tanh"),
    ("realnvp", "mlp", 9, 18, 2, 1, "swish"),
    ("realnvp", "resnet", 16, 64, 2, 2, "swish"),
    ("realnvp", "resnet", 12, 24, 2, 1, "tanh"),
]])
def test_hidden_width_below_64_runs_on_the_tensor_core_kernels(ftype, net, D, H, n_layers, launches, act, tmp_path):
    """The reference's DEFAULT conditioner width is 2 * n_inputs
    (/root/reference/src/nessai/flows/utils.py:105-165, flowmodel/utils.py:39-42), not the 64 of
    BASELINE's C2: the tcgen05 kernels take every width <= 64 (hidden units zero-padded to 64 in the
    weight images).  Checked against the float64 oracle and the generic fp32 kernel; the launch
    count proves which path ran."""
    from nessai_b200 import _lib
    from nessai_b200.flowmodel import B200FlowModel
    from test_oracle import numpy_flow

    cfg = dict(n_inputs=D, n_neurons=H, n_blocks=3, n_layers=n_layers, ftype=ftype)
    if ftype == "realnvp":
        cfg["net"] = net
        cfg["activation"] = act
    torch.manual_seed(100 * D + H)
    fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=str(tmp_path))
    fm.initialise()
    rng = np.random.default_rng(D * H)
    sd = {}
    for k, v in fm.model.state_dict().items():
        a = v.cpu().numpy()
        if a.dtype.kind == "f":
            a = a + (0.05 * rng.standard_normal(a.shape)).astype(np.float32)  # (kept well conditioned)
            if "running_var" in k:
                a = np.abs(a) + 0.5
        sd[k] = a
    fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    fm.model.eval()
    lib = _lib.load()
    n = 1000
    z = rng.normal(size=(n, D)).astype(np.float32)
    zt = torch.from_numpy(z).cuda()
    out = {}
    try:
        for tc in (1, 0):
            lib.nb200_set_tensor_core_path(tc)
            _lib.reset_launch_count()
            x, logj, logq = fm.model._inverse(zt)
            n_launch = _lib.launch_count()
            zz, flogj, logp = fm.model._forward(x)
            torch.cuda.synchronize()
            out[tc] = [t.cpu().numpy() for t in (x, logj, logq, zz, flogj, logp)]
            assert n_launch == (launches if tc else 1)
    finally:
        lib.nb200_set_tensor_core_path(1)
    nf = numpy_flow(cfg, sd)
    x64, ilj64 = nf.inverse(z.astype(np.float64))
    tol = dict(rtol=2e-4, atol=2e-4) if ftype == "nsf" else dict(rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out[1][0], x64, **tol)
    np.testing.assert_allclose(out[1][1], ilj64, **tol)
    np.testing.assert_allclose(out[1][5], nf.log_prob(out[1][0].astype(np.float64)), **tol)
    for a, b in zip(out[1], out[0]):
        assert np.isfinite(a).all()
        np.testing.assert_allclose(a, b, **tol)
    np.testing.assert_allclose(out[1][3], z, rtol=5e-4, atol=5e-4)
