"""Pins ``oracle/train_numpy.py`` (the hand-derived backward pass the CUDA training
kernels implement) against torch autograd through the reference's OWN module
tree (``nessai.flows.utils.configure_model`` on the glasflow shim), and its
clip + Adam/AdamW against stock ``torch.optim``.
"""

import copy

import numpy as np
import pytest
import torch
from conftest import load_golden, reference_or_skip

from nessai_b200.spec import FlowSpec
from oracle.train_numpy import TrainStepOracle

CASES = ["c2_realnvp_mlp", "c2_realnvp_resnet", "d5_realnvp_perm_tanh", "d4_realnvp_additive_silu", "c1_realnvp_2d",
         "d6_nsf", "d8_maf", "d5_realnvp_mvn"]


def _reference_model(cfg, sd):
    reference_or_skip()
    from nessai.flows.utils import configure_model

    model = configure_model(copy.deepcopy(cfg))
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model.double()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("weighted", [False, True])
def test_loss_and_grad_match_reference_autograd(name, weighted):
    g, cfg, sd = load_golden(name)
    model = _reference_model(cfg, sd)
    model.train()
    spec = FlowSpec(cfg)
    theta = np.zeros(spec.n_theta, dtype=np.float32)
    ints = {}
    spec.load_state_dict_numpy(sd, theta, ints)
    theta = theta.astype(np.float64)
    rng = np.random.default_rng(5)
    x = np.asarray(g["x"], dtype=np.float64)[:301]
    w = rng.uniform(0.2, 2.0, size=len(x)) if weighted else None

    xt = torch.from_numpy(x)
    lp = model.log_prob(xt)
    if weighted:
        wt = torch.from_numpy(w)
        loss_t = -torch.sum(lp * wt) / torch.sum(wt)
    else:
        loss_t = -lp.mean()
    loss_t.backward()

    oracle = TrainStepOracle(spec, ints)
    loss, grad = oracle.loss_and_grad(theta, x, weights=w)
    assert abs(loss - float(loss_t)) < 1e-9 * max(1.0, abs(loss))
    named = dict(model.named_parameters())
    checked = 0
    for e in spec.entries:
        if e.kind != "param":
            continue
        ref = named[e.key].grad.numpy().ravel()
        ours = grad[e.offset : e.offset + e.size]
        np.testing.assert_allclose(ours, ref, rtol=1e-7, atol=1e-10, err_msg=e.key)
        checked += 1
    assert checked == len(named)
    # BatchNorm running statistics were EMA-updated like the reference's buffers
    bufs = dict(model.named_buffers())
    for e in spec.entries:
        if e.kind == "fbuf":
            np.testing.assert_allclose(
                theta[e.offset : e.offset + e.size], bufs[e.key].numpy().ravel(), rtol=1e-10, atol=1e-12,
                err_msg=e.key,
            )


@pytest.mark.parametrize("opt", ["adamw", "adam"])
def test_clip_and_adam_match_torch(opt):
    rng = np.random.default_rng(3)
    p0 = rng.normal(size=500)
    p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    if opt == "adamw":
        o = torch.optim.AdamW([p], lr=1e-3)
        kw = dict(weight_decay=1e-2, decoupled=True)
    else:
        o = torch.optim.Adam([p], lr=1e-3, weight_decay=1e-6)
        kw = dict(weight_decay=1e-6, decoupled=False)
    pn, m, v = p0.copy(), np.zeros(500), np.zeros(500)
    for t in range(1, 6):
        gr = rng.normal(size=500) * (10.0 if t % 2 else 0.01)
        p.grad = torch.from_numpy(gr.copy())
        torch.nn.utils.clip_grad_norm_([p], 5.0)
        o.step()
        gn = gr.copy()
        TrainStepOracle.clip_(gn, 5.0)
        TrainStepOracle.adam_step_(pn, gn, m, v, t, 1e-3, **kw)
        np.testing.assert_allclose(pn, p.detach().numpy(), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("seed", range(10))
def test_loss_and_grad_match_reference_autograd_over_the_config_space(seed):
    """The checker of the GPU training sweep (tests/test_gpu_fuzz.py) pinned over the same space:
    float64 autograd through the reference's own module tree for random configurations (features,
    conditioner width / depth / type, linear transform, BatchNorm, activation, coupling type), ragged
    and weighted batches."""
    reference_or_skip()
    from nessai.flowmodel.utils import update_flow_config
    from nessai.flows.utils import configure_model
    from test_gpu_fuzz import draw_config

    cfg = draw_config(400 + seed)
    torch.manual_seed(seed)
    full = update_flow_config(dict(cfg))
    model = configure_model(copy.deepcopy(full))
    rng = np.random.default_rng(seed)
    sd = {}
    for k, v in model.state_dict().items():
        a = v.numpy().copy()
        if a.dtype.kind == "f" and not k.endswith(".mask"):
            a = a + (0.04 * rng.standard_normal(a.shape)).astype(np.float32)
            if "running_var" in k:
                a = np.abs(a) + 0.5
        sd[k] = a
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.double()
    model.train()
    spec = FlowSpec(dict(cfg))
    theta = np.zeros(spec.n_theta, dtype=np.float32)
    ints = {}
    spec.load_state_dict_numpy(sd, theta, ints)
    theta = theta.astype(np.float64)
    D = cfg["n_inputs"]
    x = 1.2 * rng.standard_normal((int(rng.integers(40, 400)), D)) + 0.3
    w = rng.uniform(0.2, 2.0, size=len(x)) if seed % 2 else None
    lp = model.log_prob(torch.from_numpy(x))
    loss_t = -lp.mean() if w is None else -torch.sum(lp * torch.from_numpy(w)) / float(np.sum(w))
    loss_t.backward()
    loss, grad = TrainStepOracle(spec, ints).loss_and_grad(theta, x, weights=w)
    assert abs(loss - float(loss_t)) < 1e-9 * max(1.0, abs(loss)), cfg
    named = dict(model.named_parameters())
    for e in spec.entries:
        if e.kind == "param":
            np.testing.assert_allclose(grad[e.offset : e.offset + e.size], named[e.key].grad.numpy().ravel(), rtol=1e-6,
                                       atol=1e-9, err_msg=f"{cfg} {e.key}")
