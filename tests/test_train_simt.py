"""One optimisation step of the fused training kernels (csrc/train.cuh: FWD -> LOSS -> BWD ->
REDUCE -> ADAM, the CUDA source the product runs) on the CPU under the SIMT shim of
tests/_hostcheck, against the float64 training oracle (oracle/train_numpy.py, itself pinned
against torch autograd through the reference's module tree): loss, every parameter gradient,
BatchNorm running statistics, the global-norm clip and the optimiser update -- what
tests/test_gpu_train.py checks on the GPU, with the same tolerances."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
from conftest import REPO, load_golden, simt_or_skip

from nessai_b200.spec import FlowSpec
from nessai_b200.train_plan import build_train_plan, param_mask


@pytest.fixture(scope="module")
def simt_train(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d = os.path.join(REPO, "tests", "_hostcheck")
    out = tmp_path_factory.mktemp("simt") / "libtrain_simt.so"
    res = subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", f"-I{d}/fake_cuda", "-o", str(out),
                          os.path.join(d, "train_simt.cpp")], capture_output=True, text=True)
    if res.returncode != 0:
        if "barrier" in res.stderr:
            pytest.skip("this g++ has no <barrier>")
        raise RuntimeError(res.stderr)
    lib = C.CDLL(str(out))
    lib.simt_train_step.restype = C.c_int
    lib.simt_train_step.argtypes = ([C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6
                                    + [C.c_int, C.c_int] + [C.c_double] * 6 + [C.c_int64] + [C.c_void_p] * 4 + [C.c_int])
    lib.simt_eval_loss.restype = C.c_int
    lib.simt_eval_loss.argtypes = ([C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4
                                   + [C.c_int] + [C.c_void_p] * 3 + [C.c_int])
    return simt_or_skip(lib, 512)


def setup(name):
    g, cfg, sd = load_golden(name)
    spec = FlowSpec(cfg)
    theta = np.zeros(spec.n_theta, np.float32)
    ints = {}
    spec.load_state_dict_numpy(sd, theta, ints)
    return g, spec, theta, ints


def step(lib, spec, ints, theta, x, w=None, opt=-1, lr=0.0, b1=0.0, b2=0.0, eps=0.0, wd=0.0, clip=0.0, step0=0,
         m=None, v=None):
    plan, itab, red = build_train_plan(spec, ints)
    n = spec.n_params
    theta_p, theta_b = theta[:n].copy(), theta[n:].copy()
    pm = None
    segs = param_mask(spec)
    if segs:
        pm = np.ones(n, dtype=np.float32)
        for w_off, m_off, size in segs:
            pm[w_off : w_off + size] = theta_b[m_off : m_off + size]
    m = np.zeros(n, np.float32) if m is None else m
    v = np.zeros(n, np.float32) if v is None else v
    loss, info, grad = np.zeros(1, np.float32), np.zeros(2, np.float32), np.zeros(n, np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    rc = lib.simt_train_step(plan.ctypes.data, int(plan.size), itab.ctypes.data, int(itab.size), red.ctypes.data,
                             int(red.size), theta_p.ctypes.data, ptr(theta_b) if theta_b.size else None, m.ctypes.data,
                             v.ctypes.data, x.ctypes.data, ptr(w), len(x), opt, lr, b1, b2, eps, wd, clip, step0, ptr(pm),
                             loss.ctypes.data, info.ctypes.data, grad.ctypes.data, 148)
    assert rc == 0
    return float(loss[0]), grad.astype(np.float64), info, np.concatenate([theta_p, theta_b]), m, v


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# (row counts are kept small: a CTA of the row kernels is 512 OS threads under the shim; ragged
# tiles of 16 rows and several CTAs are still covered)
@pytest.mark.parametrize("name,n_rows,weighted", [("c2_realnvp_mlp", 70, False), ("c2_realnvp_resnet", 19, True),
                                                  ("d5_realnvp_perm_tanh", 100, False),
                                                  ("d4_realnvp_additive_silu", 64, True), ("c1_realnvp_2d", 33, False),
                                                  ("d6_nsf", 50, False), ("d8_maf", 21, True),
                                                  ("d5_realnvp_mvn", 40, False)])
def test_cuda_training_step_matches_oracle(simt_train, name, n_rows, weighted):
    from oracle.train_numpy import TrainStepOracle

    g, spec, theta, ints = setup(name)
    x = np.asarray(g["train_data"], dtype=np.float64)[:n_rows].astype(np.float32)
    w = np.random.default_rng(11).uniform(0.2, 2.0, size=len(x)).astype(np.float32) if weighted else None
    theta64 = theta.astype(np.float64)
    loss64, grad64 = TrainStepOracle(spec, ints).loss_and_grad(theta64, x.astype(np.float64), weights=w)
    loss, grad, info, theta_after, _, _ = step(simt_train, spec, ints, theta, x, w)
    assert abs(loss - loss64) < 2e-5 * max(1.0, abs(loss64))
    assert np.isfinite(grad).all()
    for e in spec.entries:
        if e.kind != "param":
            continue
        a, b = grad[e.offset : e.offset + e.size], grad64[e.offset : e.offset + e.size]
        assert np.linalg.norm(a - b) <= 2e-4 * np.linalg.norm(b) + 2e-6, (e.key, rel_err(a, b))
    assert rel_err(grad, grad64) < 5e-5
    assert abs(float(info[1]) - np.linalg.norm(grad64)) < 1e-4 * np.linalg.norm(grad64)
    for e in spec.entries:  # BatchNorm running statistics (EMA of the batch statistics)
        if e.kind == "fbuf":
            np.testing.assert_allclose(theta_after[e.offset : e.offset + e.size], theta64[e.offset : e.offset + e.size],
                                       rtol=2e-5, atol=2e-6, err_msg=e.key)


@pytest.mark.parametrize("opt", ["adamw", "adam", "sgd"])
def test_cuda_clip_and_optimiser_steps_match_oracle(simt_train, opt):
    """Three consecutive clip + optimiser steps (flowmodel/base.py:439-445, torch's AdamW / Adam
    with L2 decay / SGD as stock torch defines them) of the kernels against the oracle's."""
    from oracle.train_numpy import TrainStepOracle

    g, spec, theta, ints = setup("c1_realnvp_2d")
    P, lr = spec.n_params, 3e-3
    rng = np.random.default_rng(2)
    xs = [rng.normal(size=(40, spec.D)).astype(np.float32) * 1.5 for _ in range(3)]
    theta64 = theta.astype(np.float64)
    oracle = TrainStepOracle(spec, ints)
    m64, v64 = np.zeros(P), np.zeros(P)
    kw = dict(adamw=dict(weight_decay=1e-2, decoupled=True), adam=dict(weight_decay=1e-6, decoupled=False))
    kind = dict(adamw=0, adam=1, sgd=2)[opt]
    wd = dict(adamw=1e-2, adam=1e-6, sgd=0.0)[opt]
    cur, m, v = theta.copy(), np.zeros(P, np.float32), np.zeros(P, np.float32)
    g_min, g_scale = np.full(P, np.inf), 0.0
    for t, xb in enumerate(xs, start=1):
        loss64, grad64 = oracle.loss_and_grad(theta64, xb.astype(np.float64))
        g_min = np.minimum(g_min, np.abs(grad64))
        g_scale = max(g_scale, float(np.abs(grad64).max()))
        TrainStepOracle.clip_(grad64, 0.5)  # (a clip that bites: |g| is of order one here)
        if opt == "sgd":
            theta64[:P] -= lr * grad64
        else:
            TrainStepOracle.adam_step_(theta64[:P], grad64, m64, v64, t, lr, **kw[opt])
        loss, grad, info, cur, m, v = step(simt_train, spec, ints, cur, xb, opt=kind, lr=lr, b1=0.9, b2=0.999, eps=1e-8,
                                           wd=wd, clip=0.5, step0=t - 1, m=m, v=v)
        assert abs(loss - loss64) < 1e-4 * max(1.0, abs(loss64))
    du, dv = cur[:P].astype(np.float64) - theta[:P], theta64[:P] - theta[:P]
    if opt == "sgd":
        assert rel_err(du, dv) < 1e-4
    else:
        resolved = g_min > 1e-3 * g_scale  # Adam's m / sqrt(v) amplifies fp32 rounding where g ~ 0
        assert resolved.mean() > 0.2
        assert rel_err(du[resolved], dv[resolved]) < 2e-3
        assert np.abs(du - dv).max() < 2 * 3 * lr
    tol = dict(rtol=1e-4, atol=1e-5) if opt == "sgd" else dict(rtol=5e-3, atol=3e-3)
    np.testing.assert_allclose(cur[P:], theta64[P:], **tol)


@pytest.mark.parametrize("name", ["c2_realnvp_resnet", "d5_realnvp_perm_tanh", "d6_nsf", "d8_maf", "d5_realnvp_mvn"])
def test_cuda_validation_loss_matches_reference_golden(simt_train, name):
    """tr_eval_kernel (the validation loss of FlowModel._validate: eval mode, running statistics,
    straight from the UNFOLDED parameters) against the reference's golden log-probabilities."""
    g, spec, theta, ints = setup(name)
    plan, itab, red = build_train_plan(spec, ints)
    n = spec.n_params
    theta_p, theta_b = theta[:n].copy(), theta[n:].copy()
    pm = None
    segs = param_mask(spec)
    if segs:
        pm = np.ones(n, dtype=np.float32)
        for w_off, m_off, size in segs:
            pm[w_off : w_off + size] = theta_b[m_off : m_off + size]
    rows = 40
    x = np.ascontiguousarray(g["x"][:rows], dtype=np.float32)
    w = np.random.default_rng(1).uniform(0.5, 1.5, rows).astype(np.float32)
    loss, logp = np.zeros(1, np.float32), np.zeros(rows, np.float32)
    rc = simt_train.simt_eval_loss(plan.ctypes.data, int(plan.size), itab.ctypes.data, int(itab.size), red.ctypes.data,
                                   int(red.size), theta_p.ctypes.data, theta_b.ctypes.data if theta_b.size else None,
                                   x.ctypes.data, w.ctypes.data, rows, None if pm is None else pm.ctypes.data,
                                   loss.ctypes.data, logp.ctypes.data, 148)
    assert rc == 0
    tol = 5e-4 if "nsf" in name else 1e-4
    ref = np.asarray(g["fwd_logprob64"][:rows])
    np.testing.assert_allclose(logp, ref, rtol=tol, atol=tol)
    np.testing.assert_allclose(loss[0], -np.sum(w * ref) / np.sum(w), rtol=tol, atol=tol)  # base.py:404-407
