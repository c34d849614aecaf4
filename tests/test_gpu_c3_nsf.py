"""Config C3 (BASELINE.json configs[2]): 32-D neural-spline flow, 6 coupling layers,
ResidualNet(64) conditioner, 8 bins, linear tails +-5 -- at the configuration's own
size.  There are no reference-trained weights at this size (training an NSF is the
reference's torch autograd; the fixture d6_nsf covers a reference-trained small NSF),
so the flow is randomly initialised the way ``configure_model`` does (bit-identical
init is pinned in tests/test_spec.py) with its weights scaled up so that the splines
are far from the identity, and the kernels are checked against the float64 oracle on
a sample and through size-independent properties on the full 2e6 rows."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(n_inputs=32, ftype="nsf", n_blocks=6, n_layers=2, n_neurons=64)


def weights(seed=3):
    """Same construction as tests/golden/make_c3_reference.py::weights."""
    from nessai_b200.spec import FlowSpec

    torch.manual_seed(seed)
    spec = FlowSpec(dict(CFG))
    theta, ints = spec.init_state()
    sd = spec.state_dict_numpy(theta, ints)
    rng = np.random.default_rng(seed)
    for k, v in sd.items():
        if k.endswith("final_layer.weight"):
            sd[k] = (v + 0.3 * rng.standard_normal(v.shape)).astype(np.float32)  # non-trivial splines
        elif k.endswith("final_layer.bias"):
            sd[k] = (v + 0.5 * rng.standard_normal(v.shape)).astype(np.float32)
    return sd


def make(tmp_path, seed=3):
    from nessai_b200.flowmodel import B200FlowModel

    sd = weights(seed)
    fm = B200FlowModel(flow_config=dict(CFG), training_config=dict(device_tag="cuda:0"), output=str(tmp_path))
    fm.initialise()
    fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    fm.model.eval()
    return fm, sd


def test_c3_matches_reference_within_its_own_fp32_error(tmp_path):
    """Golden vectors of the unmodified reference (torch fp32, CPU) and of the float64
    oracle at C3.  A 6-layer spline flow amplifies fp32 rounding (the reference's own
    fp32 outputs differ from float64 by up to 1e-2 in log|J|), so the bar is: our
    error against float64 stays within 4x the reference's own fp32 error, quantile
    by quantile (the tcgen05 path carries 22 significand bits through its GEMMs --
    split fp16 -- against fp32's 24; measured 2.3-3x), and typical differences to the
    reference are at the 1e-5 level.  The generic fp32 kernel is held to 2x."""
    import os

    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "c3_nsf_reference.npz"))
    from nessai_b200 import _lib

    fm, sd = make(tmp_path)
    assert abs(sum(float(np.abs(v).sum()) for v in sd.values()) - float(g["w_checksum"])) < 1e-3
    z = g["z"].astype(np.float64)

    def check(ours, ref32, f64, name, factor):
        e_ours, e_ref = np.abs(ours - f64), np.abs(ref32 - f64)
        for q in (0.5, 0.99, 0.999):
            assert np.quantile(e_ours, q) <= factor * np.quantile(e_ref, q) + 1e-7, (name, q, factor)
        # the single worst element sits on an almost-vertical spline segment (the reference's own
        # worst fp32 error there is 3e-2): order of magnitude only
        assert e_ours.max() <= 10 * e_ref.max() + 1e-3, (name, e_ours.max(), e_ref.max())
        assert np.median(np.abs(ours - ref32)) < 1e-4, name

    lib = _lib.load()
    try:
        for tc, factor, launches in ((1, 4.0, 6), (0, 2.0, 1)):
            lib.nb200_set_tensor_core_path(tc)
            _lib.reset_launch_count()
            x, logj = fm.inverse(z)
            assert _lib.launch_count() == launches  # tcgen05 path: one launch per coupling layer
            zf, logp = fm.forward_and_log_prob(g["inv_x"].astype(np.float64))
            check(x, g["inv_x"], g["inv_x64"], "x", factor)
            check(logj, g["inv_logj"], g["inv_logj64"], "logj", factor)
            check(zf, g["fwd_z"], g["fwd_z64"], "fwd z", factor)
            check(logp, g["fwd_logprob"], g["fwd_logprob64"], "logp", factor)
            # north_star tolerance on the bulk: rtol 1e-4 (+ atol) for >= 99% of the values
            for a, b in ((x, g["inv_x"]), (logj, g["inv_logj"]), (logp, g["fwd_logprob"])):
                assert (np.abs(a - b) <= 1e-4 * np.abs(b) + 5e-4).mean() >= 0.99
    finally:
        lib.nb200_set_tensor_core_path(1)


def test_c3_full_size_properties(tmp_path):
    """2e6 rows: inverse(forward) round trip, log|J| antisymmetry, sample_and_log_prob
    consistency (tests/test_flows/test_included_flows.py:129-154), determinism."""
    fm, sd = make(tmp_path)
    n = 2_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.randn(n, 32, device="cuda", generator=g)
    x, lj, lq = fm.model._inverse(z)
    zr, flj, lp = fm.model._forward(x)
    ok = torch.isfinite(lq) & torch.isfinite(lp)
    assert float(ok.float().mean()) > 0.999
    # a 6-layer random spline flow is ill-conditioned in its tails (see the golden
    # test above): bound the bulk tightly and the worst of 2e6 rows loosely
    for err, name in (((zr - z).abs().amax(1), "z"), ((lj + flj).abs(), "logj"), ((lp - lq).abs(), "logq")):
        e = err[ok]
        q = torch.quantile(e[:1_000_000].float(), torch.tensor([0.5, 0.999], device="cuda"))
        assert float(q[0]) < 5e-5, (name, q)
        assert float(q[1]) < 2e-3, (name, q)
        assert float(e.max()) < 2e-2, (name, float(e.max()))
    x2, lj2, lq2 = fm.model._inverse(z)
    assert torch.equal(x, x2) and torch.equal(lq[ok], lq2[ok])


def test_c3_training_gradient_and_descent(tmp_path):
    """FlowModel.train on the C3 spline flow runs on the fused kernels: every parameter
    gradient against the float64 training oracle (rational-quadratic spline backward,
    oracle/train_numpy.py, itself pinned against autograd through the reference's
    NeuralSplineFlow), then a short training run must reduce the loss."""
    from nessai_b200.flowmodel import B200FlowModel
    from oracle.train_numpy import TrainStepOracle

    fm, sd = make(tmp_path)
    spec = fm.model.spec
    rng = np.random.default_rng(4)
    x = np.clip(rng.normal(size=(333, 32)) * 1.2, -6, 6).astype(np.float32)
    theta64 = fm.model.theta_numpy().astype(np.float64)
    loss64, grad64 = TrainStepOracle(spec, fm.model.ints).loss_and_grad(theta64, x.astype(np.float64))
    loss, grad, info = fm._trainer().loss_and_grad(torch.from_numpy(x).cuda())
    assert abs(float(loss) - loss64) < 1e-4 * abs(loss64)
    grad = grad.cpu().numpy().astype(np.float64)
    for e in spec.entries:
        if e.kind == "param":
            a, b = grad[e.offset : e.offset + e.size], grad64[e.offset : e.offset + e.size]
            assert np.linalg.norm(a - b) <= 1e-3 * np.linalg.norm(b) + 1e-5, e.key
    assert np.linalg.norm(grad - grad64) < 2e-4 * np.linalg.norm(grad64)

    torch.manual_seed(1)
    fm2 = B200FlowModel(flow_config=dict(CFG), output=str(tmp_path / "t"), rng=np.random.default_rng(1),
                        training_config=dict(device_tag="cuda:0", max_epochs=30, patience=30, batch_size=1000, lr=3e-3))
    fm2.initialise()
    data = rng.normal(size=(4000, 32))
    data[:, 1::2] = 0.5 * data[:, 0::2] ** 2 + 0.3 * data[:, 1::2]  # curved, non-Gaussian
    data = (data - data.mean(0)) / data.std(0)
    hist = fm2.train(data, plot=False)
    assert np.isfinite(hist["loss"]).all() and np.isfinite(hist["val_loss"]).all()
    assert hist["loss"][-1] < hist["loss"][0] - 1.0
    x_s, lq = fm2.sample_and_log_prob(N=2000)
    ok = np.isfinite(lq)
    assert ok.mean() > 0.99
    np.testing.assert_allclose(fm2.log_prob(x_s)[ok], lq[ok], rtol=2e-3, atol=2e-2)
