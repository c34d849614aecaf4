"""Host logic: flat layout, initialisation, folding and program encoding."""

import numpy as np
import pytest
from _program_interp import run_program
from conftest import reference_or_skip

from nessai_b200.spec import FlowSpec


def load_spec(cfg, sd):
    sp = FlowSpec(cfg)
    theta = np.zeros(sp.n_theta, np.float32)
    ints = {}
    sp.load_state_dict_numpy(sd, theta, ints)
    return sp, theta, ints


def test_state_dict_layout_matches_reference_keys(golden):
    name, g, cfg, sd = golden
    sp, theta, ints = load_spec(cfg, sd)
    mine = sp.state_dict_numpy(theta, ints)
    assert list(mine) == list(sd)  # same keys, same order
    for k in sd:
        assert mine[k].shape == sd[k].shape
        np.testing.assert_array_equal(mine[k], sd[k])


def test_folded_program_matches_golden(golden):
    name, g, cfg, sd = golden
    if "nsf" in name:
        pytest.skip("spline op is checked on the GPU (numpy interpreter covers affine couplings)")
    sp, theta, ints = load_spec(cfg, sd)
    ff = sp.fold(theta, ints)
    z, lj = run_program(ff.program(False), g["x"])
    x, ilj = run_program(ff.program(True), g["z"])
    np.testing.assert_allclose(z, g["fwd_z64"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(lj, g["fwd_logj64"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(x, g["inv_x64"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(ilj, g["inv_logj64"], atol=2e-5, rtol=1e-5)


def test_program_fp32_within_stated_tolerance(golden):
    """fp32 evaluation of the folded program vs the reference's fp32 outputs:
    log_prob / log|J| rtol 1e-4 (BASELINE.json north_star)."""
    name, g, cfg, sd = golden
    if "nsf" in name:
        pytest.skip("spline op is checked on the GPU")
    sp, theta, ints = load_spec(cfg, sd)
    ff = sp.fold(theta, ints)
    z, lj = run_program(ff.program(False), g["x"], np.float32)
    var = sp.base_var  # N(0, var I) base distribution (1 except for the "mvn" fixture)
    lp = -(0.5 / var) * (z.astype(np.float64) ** 2).sum(1) - 0.5 * sp.D * np.log(2 * np.pi * var) + lj
    np.testing.assert_allclose(lp, g["fwd_logprob"], rtol=1e-4, atol=1e-4)


@pytest.mark.reference
@pytest.mark.parametrize(
    "cfg",
    [
        dict(n_inputs=16, n_neurons=64, n_blocks=4, n_layers=2, ftype="realnvp", net="mlp"),
        dict(n_inputs=3, n_neurons=None, n_blocks=2, n_layers=2, ftype="realnvp"),
        dict(n_inputs=6, n_neurons=16, n_blocks=3, n_layers=2, ftype="nsf"),
        dict(n_inputs=5, n_neurons=8, n_blocks=2, n_layers=1, ftype="realnvp",
             linear_transform="permutation", net="mlp"),
    ],
)
def test_init_is_bit_identical_to_configure_model(cfg):
    """Same torch seed -> same initial state_dict as the reference's
    configure_model (flows/utils.py:208-246)."""
    reference_or_skip()
    import torch
    from nessai.flowmodel.utils import update_flow_config
    from nessai.flows import configure_model

    torch.manual_seed(42)
    ref = configure_model(update_flow_config(cfg)).state_dict()
    sp = FlowSpec(cfg)
    torch.manual_seed(42)
    theta, ints = sp.init_state()
    mine = sp.state_dict_numpy(theta, ints)
    assert list(mine) == list(ref)
    for k, v in ref.items():
        np.testing.assert_array_equal(mine[k], v.numpy())


@pytest.mark.reference
def test_reset_weights_matches_reference():
    reference_or_skip()
    import torch
    from nessai.flows import configure_model
    from nessai.flows.utils import reset_weights

    cfg = dict(n_inputs=4, n_neurons=8, n_blocks=2, n_layers=2, ftype="realnvp")
    torch.manual_seed(0)
    m = configure_model(cfg)
    sp = FlowSpec(cfg)
    torch.manual_seed(0)
    theta, ints = sp.init_state()
    torch.manual_seed(5)
    m.apply(reset_weights)
    torch.manual_seed(5)
    sp.reset_weights(theta)
    mine = sp.state_dict_numpy(theta, ints)
    for k, v in m.state_dict().items():
        np.testing.assert_array_equal(mine[k], v.numpy())


def test_unsupported_options_fail_loudly():
    with pytest.raises(NotImplementedError):
        FlowSpec(dict(n_inputs=4, ftype="maf", use_random_masks=True))
    with pytest.raises(NotImplementedError):
        FlowSpec(dict(n_inputs=4, ftype="glow"))
    with pytest.raises(NotImplementedError):
        FlowSpec(dict(n_inputs=4, ftype="realnvp", linear_transform="svd"))
    with pytest.raises(NotImplementedError):
        FlowSpec(dict(n_inputs=4, ftype="realnvp", distribution="lars"))
    with pytest.raises(ValueError):
        FlowSpec(dict(n_inputs=1, ftype="realnvp"))
    with pytest.raises(TypeError):
        FlowSpec(dict(n_inputs=4.0, ftype="realnvp"))


def test_n_neurons_auto():
    assert FlowSpec(dict(n_inputs=5, ftype="realnvp")).H == 10
    assert FlowSpec(dict(n_inputs=5, n_neurons="equal", ftype="realnvp")).H == 5


def test_train_plan_packs_every_golden_realnvp():
    """The packed TrPlan (csrc/train.cuh) covers every parameter exactly once:
    conditioner weights / biases and LU biases through reduce_idx, LU triangles and
    BatchNorm parameters through their layer records."""
    from conftest import GOLDEN_NAMES, load_golden

    from nessai_b200.spec import FlowSpec
    from nessai_b200.train_plan import TR_LAYER_INTS, TR_PLAN_INTS, TrainPlanUnsupported, build_train_plan

    for name in GOLDEN_NAMES:
        g, cfg, sd = load_golden(name)
        spec = FlowSpec(cfg)
        theta = np.zeros(spec.n_theta, dtype=np.float32)
        ints = {}
        spec.load_state_dict_numpy(sd, theta, ints)
        plan, itab, red = build_train_plan(spec, ints)
        assert plan.size == TR_PLAN_INTS and plan[0] == spec.D and plan[1] == spec.L
        covered = np.zeros(spec.n_params, dtype=int)
        covered[red] += 1
        D = spec.D
        ntri = D * (D - 1) // 2
        for l in range(spec.L):
            row = plan[20 + l * TR_LAYER_INTS : 20 + (l + 1) * TR_LAYER_INTS]
            if row[1] >= 0:
                covered[row[2] : row[2] + ntri] += 1
                covered[row[3] : row[3] + ntri] += 1
                covered[row[4] : row[4] + D] += 1
            if row[5] >= 0:
                covered[row[5] : row[5] + D] += 1
                covered[row[6] : row[6] + D] += 1
            n_buf = row[14]
            dims = row[20 : 20 + n_buf]
            assert row[15] == 2 * D + dims[1:].sum()
            assert dims[0] == row[11] and dims[-1] == row[12] * spec.coupling_multiplier
            assert row[11] == (D if spec.ftype == "maf" else len(spec.layers[l].identity))
        assert (covered == 1).all()


@pytest.mark.reference
@pytest.mark.parametrize("mask", [[1, 1, -1], [1, 1, 1, -1, -1], [[1, -1, 1], [-1, 1, 1]]])
def test_custom_mask_of_the_augmented_proposal(mask):
    """``AugmentedFlowProposal.update_flow_config`` (proposal/augmented.py:91-96) passes
    ``mask = ones(D); mask[-augment_dims:] = -1`` (realnvp.py:114-131 alternates its sign per
    layer; a 2-D mask gives every layer its own): same initial state_dict as the reference,
    the folded program reproduces the reference module in both directions, and the training
    plan packs it."""
    reference_or_skip()
    import torch
    from nessai.flowmodel.utils import update_flow_config
    from nessai.flows import configure_model

    from nessai_b200.train_plan import build_train_plan

    D = np.shape(mask)[-1]
    cfg = dict(n_inputs=D, n_neurons=8, n_blocks=2, n_layers=2, ftype="realnvp", mask=np.array(mask))
    torch.manual_seed(7)
    ref = configure_model(update_flow_config(dict(cfg)))
    ref.eval()
    sp = FlowSpec(cfg)
    torch.manual_seed(7)
    theta, ints = sp.init_state()
    mine = sp.state_dict_numpy(theta, ints)
    sd = ref.state_dict()
    assert list(mine) == list(sd)
    for k, v in sd.items():
        np.testing.assert_array_equal(mine[k], v.numpy())
    # perturb every parameter so that the couplings, LU and BatchNorm all act
    rng = np.random.default_rng(1)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(torch.from_numpy(0.2 * rng.standard_normal(tuple(p.shape))).to(p.dtype))
        for name, b in ref.named_buffers():
            if name.endswith("running_var"):
                b.mul_(1.7)
            elif name.endswith("running_mean"):
                b.add_(0.3)
    sd = {k: v.numpy() for k, v in ref.state_dict().items()}
    sp.load_state_dict_numpy(sd, theta, ints)
    ff = sp.fold(theta, ints)
    x = rng.standard_normal((64, D))
    with torch.no_grad():
        z_ref, lj_ref = ref.forward(torch.from_numpy(x).float())
        x_ref, ilj_ref = ref.inverse(torch.from_numpy(x).float())
    z, lj = run_program(ff.program(False), x)
    xi, ilj = run_program(ff.program(True), x)
    np.testing.assert_allclose(z, z_ref.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(lj, lj_ref.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(xi, x_ref.numpy(), rtol=2e-4, atol=5e-5)
    np.testing.assert_allclose(ilj, ilj_ref.numpy(), rtol=2e-4, atol=2e-5)
    plan, itab, red = build_train_plan(sp, ints)
    assert plan[0] == D and plan[1] == 2


def test_interchange_with_real_glasflow():
    """Everything above pins the nflows boundary against ``oracle/shims/glasflow`` -- a restatement
    written in this repository -- because the real ``glasflow`` is absent from this image.  Where it
    IS installed, this test checks the claims directly: ``configure_model`` on real glasflow gives
    the state_dict layout / initial values ``FlowSpec`` lays out, weights load in both directions,
    and forward / inverse / log_prob of the float64 oracle agree with the real module."""
    glasflow = pytest.importorskip("glasflow")
    if "oracle-shim" in getattr(glasflow, "__version__", ""):
        pytest.skip("only the restated glasflow shim is importable here (real glasflow not installed)")
    import torch
    from nessai.flows.utils import configure_model

    from nessai_b200.spec import FlowSpec
    from oracle.flow_numpy import NumpyFlow

    for cfg in (dict(n_inputs=6, n_neurons=16, n_blocks=3, n_layers=2, ftype="realnvp"),
                dict(n_inputs=6, n_neurons=16, n_blocks=3, n_layers=2, ftype="realnvp", net="mlp"),
                dict(n_inputs=6, n_neurons=16, n_blocks=2, n_layers=2, ftype="nsf"),
                dict(n_inputs=6, n_neurons=16, n_blocks=2, n_layers=2, ftype="maf")):
        torch.manual_seed(7)
        model = configure_model(dict(cfg)).eval()
        torch.manual_seed(7)
        spec = FlowSpec(dict(cfg))
        theta, ints = spec.init_state()
        ours = spec.state_dict_numpy(theta, ints)
        ref = {k: v.detach().numpy() for k, v in model.state_dict().items()}
        assert list(ours) == list(ref)
        for k in ref:
            np.testing.assert_array_equal(ours[k], ref[k], err_msg=k)
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in ours.items()})  # our layout loads
        x = torch.randn(64, 6)
        with torch.no_grad():
            z, lj = model.forward(x)
            lp = model.log_prob(x)
        nf = NumpyFlow(ref, ftype=cfg["ftype"], net=cfg.get("net", "resnet"), hidden_features=16)
        z64, lj64 = nf.forward(x.numpy())
        np.testing.assert_allclose(z64, z.numpy(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(lj64, lj.numpy(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(nf.log_prob(x.numpy()), lp.numpy(), rtol=1e-4, atol=1e-4)
