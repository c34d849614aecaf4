"""Parity of the fused training kernels (csrc/train.cuh, through the C ABI) with
the float64 training oracle (oracle/train_numpy.py, itself pinned against autograd
through the reference's module tree): loss, every parameter gradient, BatchNorm
running statistics, clipping and the optimiser update; plus the eval-mode loss
against the inference kernels, and the on-device autograd cross-check.
"""

import numpy as np
import pytest
import torch
from conftest import load_golden

pytestmark = pytest.mark.gpu

CASES = ["c2_realnvp_mlp", "c2_realnvp_resnet", "d5_realnvp_perm_tanh", "d4_realnvp_additive_silu", "c1_realnvp_2d",
         "d6_nsf", "d8_maf", "d5_realnvp_mvn"]


def make_model(cfg, sd, tmp_path, **training):
    from nessai_b200.flowmodel import B200FlowModel

    tc = dict(device_tag="cuda:0")
    tc.update(training)
    fm = B200FlowModel(flow_config=cfg, training_config=tc, output=str(tmp_path))
    fm.initialise()
    fm.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return fm


def oracle_for(fm):
    from oracle.train_numpy import TrainStepOracle

    return TrainStepOracle(fm.model.spec, fm.model.ints)


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("n_rows,weighted", [(301, False), (1000, False), (37, True)])
def test_loss_and_gradient_match_oracle(name, n_rows, weighted, tmp_path):
    g, cfg, sd = load_golden(name)
    fm = make_model(cfg, sd, tmp_path)
    spec = fm.model.spec
    x = np.asarray(g["train_data"], dtype=np.float64)[:n_rows].astype(np.float32)
    rng = np.random.default_rng(11)
    w = rng.uniform(0.2, 2.0, size=len(x)).astype(np.float32) if weighted else None
    theta64 = fm.model.theta_numpy().astype(np.float64)
    loss64, grad64 = oracle_for(fm).loss_and_grad(theta64, x.astype(np.float64), weights=w)

    xt = torch.from_numpy(x).cuda()
    wt = None if w is None else torch.from_numpy(w).cuda()
    loss, grad, info = fm._trainer().loss_and_grad(xt, wt)
    torch.cuda.synchronize()
    assert abs(float(loss) - loss64) < 2e-5 * max(1.0, abs(loss64))
    grad = grad.cpu().numpy().astype(np.float64)
    assert np.isfinite(grad).all()
    # per tensor: relative L2 error of fp32 kernels vs the float64 oracle
    for e in spec.entries:
        if e.kind != "param":
            continue
        a, b = grad[e.offset : e.offset + e.size], grad64[e.offset : e.offset + e.size]
        assert np.linalg.norm(a - b) <= 2e-4 * np.linalg.norm(b) + 2e-6, (e.key, rel_err(a, b))
    assert rel_err(grad, grad64) < 5e-5
    assert abs(float(info[1]) - np.linalg.norm(grad64)) < 1e-4 * np.linalg.norm(grad64)
    # BatchNorm running statistics (EMA of the batch statistics)
    theta_after = fm.model.theta_numpy()
    for e in spec.entries:
        if e.kind == "fbuf":
            np.testing.assert_allclose(
                theta_after[e.offset : e.offset + e.size], theta64[e.offset : e.offset + e.size],
                rtol=2e-5, atol=2e-6, err_msg=e.key,
            )


@pytest.mark.parametrize("name", ["c2_realnvp_mlp", "c2_realnvp_resnet"])
@pytest.mark.parametrize("opt", ["adamw", "adam", "sgd"])
def test_epoch_of_steps_matches_oracle(name, opt, tmp_path):
    """Three batches (1000 + 1000 + 500 rows, permuted) of clip + optimiser steps."""
    from oracle.train_numpy import TrainStepOracle

    g, cfg, sd = load_golden(name)
    fm = make_model(cfg, sd, tmp_path, optimiser=opt, lr=3e-3)
    spec = fm.model.spec
    rng = np.random.default_rng(2)
    x = rng.normal(size=(2500, spec.D)).astype(np.float32) * 1.5
    perm = rng.permutation(len(x))
    theta = fm.model.theta_numpy().astype(np.float64)
    oracle = oracle_for(fm)
    P = spec.n_params
    m, v = np.zeros(P), np.zeros(P)
    kw = dict(adamw=dict(weight_decay=1e-2, decoupled=True), adam=dict(weight_decay=1e-6, decoupled=False))
    losses = []
    g_min, g_scale = np.full(P, np.inf), 0.0
    for t, i0 in enumerate(range(0, len(x), 1000), start=1):
        xb = x[perm[i0 : i0 + 1000]].astype(np.float64)
        loss, grad = oracle.loss_and_grad(theta, xb)
        losses.append(loss)
        g_min = np.minimum(g_min, np.abs(grad))
        g_scale = max(g_scale, float(np.abs(grad).max()))
        TrainStepOracle.clip_(grad, 5.0)
        if opt == "sgd":
            theta[:P] -= 3e-3 * grad
        else:
            TrainStepOracle.adam_step_(theta[:P], grad, m, v, t, 3e-3, **kw[opt])
    fm._batch_size = 1000
    total = fm._trainer().epoch(
        torch.from_numpy(x).cuda(), None, torch.from_numpy(perm).cuda(), 1000, fm._optimiser, 5.0
    )
    torch.cuda.synchronize()
    assert abs(float(total) - sum(losses)) < 1e-4 * abs(sum(losses))
    after = fm.model.theta_numpy().astype(np.float64)
    # parameters moved by ~lr per step; compare the UPDATE, not the value
    before = np.zeros(spec.n_theta)
    spec_theta = np.zeros(spec.n_theta, dtype=np.float32)
    ints = {}
    spec.load_state_dict_numpy(sd, spec_theta, ints)
    before[:] = spec_theta
    du, dv = after[:P] - before[:P], theta[:P] - before[:P]
    if opt == "sgd":
        assert rel_err(du, dv) < 1e-4, rel_err(du, dv)
    else:
        # Adam's m / sqrt(v) amplifies fp32 rounding where the gradient is ~0 (the update
        # has magnitude ~lr whatever |g| is): compare where the gradient is resolved, and
        # bound the rest by the step size
        resolved = g_min > 1e-3 * g_scale
        assert resolved.mean() > 0.2
        assert rel_err(du[resolved], dv[resolved]) < 2e-3, rel_err(du[resolved], dv[resolved])
        assert np.abs(du - dv).max() < 2 * 3 * 3e-3
    # running statistics follow the parameters: tight for SGD, within the Adam noise otherwise
    tol = dict(rtol=1e-4, atol=1e-5) if opt == "sgd" else dict(rtol=5e-3, atol=3e-3)
    np.testing.assert_allclose(after[P:], theta[P:], **tol)


@pytest.mark.parametrize("name", CASES)
def test_eval_loss_matches_inference_kernels(name, tmp_path):
    """The eval-mode pass over the UNFOLDED parameters equals the folded inference
    program (the reference pins forward_and_log_prob == forward + log_prob bit for
    bit, tests/test_flows/test_included_flows.py:114-126; here two different
    kernels, so fp32 tolerance)."""
    g, cfg, sd = load_golden(name)
    fm = make_model(cfg, sd, tmp_path)
    fm.model.eval()
    x = torch.from_numpy(np.asarray(g["x"], dtype=np.float32)).cuda()
    lp_folded = fm.model.log_prob(x).cpu().numpy()
    tr = fm._trainer()
    lp = tr.log_prob_unfolded(x).cpu().numpy()
    np.testing.assert_allclose(lp, g["fwd_logprob64"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(lp, lp_folded, rtol=1e-4, atol=1e-4)
    loss = float(tr.eval_loss(x).cpu())
    assert abs(loss + float(np.mean(g["fwd_logprob64"]))) < 1e-4 * abs(loss)


@pytest.mark.parametrize("name", ["c2_realnvp_resnet", "d5_realnvp_perm_tanh"])
def test_gradient_matches_device_autograd(name, tmp_path):
    from _eager_flow import EagerFlow

    g, cfg, sd = load_golden(name)
    fm = make_model(cfg, sd, tmp_path)
    model = fm.model
    x = torch.from_numpy(np.asarray(g["train_data"], dtype=np.float32)[:777]).cuda()
    tb = model.theta_b.clone()
    eager = EagerFlow(model.spec, model.ints, model.device)
    tp = model.theta_p.detach().clone().requires_grad_(True)
    loss_t = -eager.log_prob((tp, tb), x, training=True).mean()
    loss_t.backward()
    loss, grad, _ = fm._trainer().loss_and_grad(x)
    assert abs(float(loss) - float(loss_t)) < 2e-5 * abs(float(loss_t))
    ref = tp.grad.cpu().numpy()
    assert rel_err(grad.cpu().numpy(), ref) < 1e-4


@pytest.mark.parametrize("name,patience,annealing,weighted", [
    ("c2_realnvp_mlp", 3, False, False),
    ("c2_realnvp_resnet", 2, True, False),
    ("d6_nsf", 2, False, True),
    ("d8_maf", 50, False, False),
])
def test_device_epoch_loop_equals_host_epoch_loop(name, patience, annealing, weighted, tmp_path):
    """``FlowModel.train`` with the epoch loop on the device (persistent kernel: steps, validation
    loss, best-weights snapshot, patience -- flowmodel/base.py:620-667) against the same loop driven
    epoch by epoch from the host: same history, same stopping epoch, same final weights, and torch's
    CPU generator left where the reference's loop would leave it (one ``randperm`` per epoch run)."""
    from nessai_b200.flowmodel import B200FlowModel

    g, cfg, sd = load_golden(name)
    x = np.asarray(g["train_data"], dtype=np.float64)
    w = np.random.default_rng(5).uniform(0.5, 1.5, size=len(x)) if weighted else None
    out = {}
    for device_loop in (True, False):
        torch.manual_seed(1234)
        fm = B200FlowModel(
            flow_config=cfg, output=str(tmp_path / str(device_loop)), rng=np.random.default_rng(7),
            training_config=dict(max_epochs=40, patience=patience, batch_size=700, lr=0.01, annealing=annealing))
        fm.initialise()
        fm._device_loop = device_loop
        hist = fm.train(x, weights=w, plot=False)
        out[device_loop] = (hist, fm.model.theta_numpy().copy(), fm.model.theta_b.cpu().numpy().copy(),
                            torch.get_rng_state().clone(), float(fm._optimiser.param_groups[0]["lr"]))
    (h1, p1, b1, r1, lr1), (h0, p0, b0, r0, lr0) = out[True], out[False]
    assert len(h1["loss"]) == len(h0["loss"]) and len(h1["loss"]) >= 2
    if patience < 10:
        assert len(h1["loss"]) < 40  # the run stopped on patience, inside a chunk of epochs
    np.testing.assert_allclose(h1["loss"], h0["loss"], rtol=1e-6)
    np.testing.assert_allclose(h1["val_loss"], h0["val_loss"], rtol=1e-6)
    np.testing.assert_allclose(p1, p0, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(b1, b0, rtol=1e-5, atol=1e-7)
    assert torch.equal(r1, r0)
    assert lr1 == pytest.approx(lr0, rel=1e-12)
