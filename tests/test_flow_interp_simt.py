"""The generic fp32 flow interpreter -- csrc/flow_interp.cuh, the CUDA source, unchanged -- run
on the CPU under the SIMT shim (tests/_hostcheck) on the folded programs the product uploads
(``FlowSpec.fold(...).program(...)``, byte for byte what ``nb200_flow_set_program`` receives),
against the GOLDEN VECTORS of the unmodified reference: forward / inverse, log|J|, log-prob for
every flow family incl. the rational-quadratic spline op and the MADE flags, at the tolerance
the north star states for the GPU path (rtol 1e-4 on log_prob / log|J|)."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
from conftest import REPO, load_golden, simt_or_skip

from nessai_b200.spec import FlowSpec

RTOL, ATOL = 1e-4, 1e-4


@pytest.fixture(scope="module")
def simt_interp(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d = os.path.join(REPO, "tests", "_hostcheck")
    out = tmp_path_factory.mktemp("simt") / "libflow_interp_simt.so"
    res = subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", f"-I{d}/fake_cuda", "-o", str(out),
                          os.path.join(d, "flow_interp_simt.cpp")], capture_output=True, text=True)
    if res.returncode != 0:
        if "barrier" in res.stderr:
            pytest.skip("this g++ has no <barrier>")
        raise RuntimeError(res.stderr)
    lib = C.CDLL(str(out))
    lib.simt_flow_apply.restype = C.c_int
    lib.simt_flow_apply.argtypes = ([C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 4 + [C.c_double]
                                    + [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_double])
    return simt_or_skip(lib, 128)


def run(lib, spec, prog, rows, lp_mode, grid=2):
    ops = np.ascontiguousarray(prog.ops, dtype=np.int32)
    blob = np.ascontiguousarray(prog.blob, dtype=np.float32)
    x = np.ascontiguousarray(rows, dtype=np.float32)
    n = len(x)
    out = np.full((n, spec.D), np.nan, dtype=np.float32)
    logj = np.full(n, np.nan, dtype=np.float32)
    lp = np.full(n, np.nan, dtype=np.float32)
    rc = lib.simt_flow_apply(grid, ops.ctypes.data, int(ops.shape[0]), blob.ctypes.data, spec.D, spec.H, spec.activation,
                             int(prog.final_buf), float(prog.const_logdet), x.ctypes.data, out.ctypes.data,
                             logj.ctypes.data, lp.ctypes.data, n, lp_mode, float(spec.base_var))
    assert rc == 0
    return out.astype(np.float64), logj.astype(np.float64), lp.astype(np.float64)


def test_cuda_interpreter_matches_reference_golden_vectors(simt_interp, golden):
    name, g, cfg, sd = golden
    sp = FlowSpec(cfg)
    theta = np.zeros(sp.n_theta, np.float32)
    ints = {}
    sp.load_state_dict_numpy(sd, theta, ints)
    ff = sp.fold(theta, ints)
    # the spline flow amplifies fp32 rounding (tests/test_gpu_c3_nsf.py ties its budget to the reference's own
    # fp32 error); the other families are well inside the stated tolerance
    rtol, atol = (5e-4, 5e-4) if "nsf" in name else (RTOL, ATOL)
    n = min(len(g["x"]), 300)
    z, lj, lp = run(simt_interp, sp, ff.program(False), g["x"][:n], lp_mode=2)
    np.testing.assert_allclose(z, g["fwd_z64"][:n], rtol=rtol, atol=atol)
    np.testing.assert_allclose(lj, g["fwd_logj64"][:n], rtol=rtol, atol=atol)
    np.testing.assert_allclose(lp, g["fwd_logprob64"][:n], rtol=rtol, atol=atol)
    np.testing.assert_allclose(lp, g["fwd_logprob"][:n], rtol=rtol, atol=atol)  # the reference's own fp32 run
    x, ilj, lq = run(simt_interp, sp, ff.program(True), g["z"][:n], lp_mode=1)
    np.testing.assert_allclose(x, g["inv_x64"][:n], rtol=rtol, atol=atol)
    np.testing.assert_allclose(ilj, g["inv_logj64"][:n], rtol=rtol, atol=atol)
    np.testing.assert_allclose(lq, g["inv_logq"][:n], rtol=rtol, atol=atol)
    # the properties the reference's tests pin (test_included_flows.py:144-154): invertibility and
    # log|J_fwd| = -log|J_inv|, through the two CUDA programs
    back, blj, _ = run(simt_interp, sp, ff.program(True), z, lp_mode=1, grid=3)
    np.testing.assert_allclose(back, g["x"][:n], rtol=10 * rtol, atol=10 * atol)
    np.testing.assert_allclose(blj, -lj, rtol=10 * rtol, atol=10 * atol)


# ------------------------------------------------------------------ the fused populate turn (generic kernel)
@pytest.fixture(scope="module")
def simt_populate(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d = os.path.join(REPO, "tests", "_hostcheck")
    out = tmp_path_factory.mktemp("simt") / "libpopulate_simt.so"
    res = subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", f"-I{d}/fake_cuda", "-o", str(out),
                          os.path.join(d, "populate_draw_simt.cpp")], capture_output=True, text=True)
    if res.returncode != 0:
        if "barrier" in res.stderr:
            pytest.skip("this g++ has no <barrier>")
        raise RuntimeError(res.stderr)
    lib = C.CDLL(str(out))
    lib.simt_populate_draw.restype = C.c_int
    lib.simt_populate_draw.argtypes = ([C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 4 + [C.c_double, C.c_int64,
                                       C.c_uint64, C.c_uint64, C.c_float, C.c_float] + [C.c_void_p] * 4
                                       + [C.c_double, C.c_double] + [C.c_void_p] * 5)
    return simt_or_skip(lib, 128)


@pytest.mark.parametrize("name,sqrt_t,min_log_q", [("c2_realnvp_mlp", 1.0, None), ("c1_realnvp_2d", 1.0, None),
                                                   ("d5_realnvp_perm_tanh", 1.3, None), ("c2_realnvp_resnet", 1.0, -24.0),
                                                   ("d6_nsf", 1.0, None), ("d8_maf", 1.0, None),
                                                   ("d5_realnvp_mvn", 1.2, None)])
def test_cuda_fused_populate_turn_matches_oracle(simt_populate, name, sqrt_t, min_log_q):
    """One fused turn of the generic kernel -- the CUDA sources of the Philox draw, the flow
    interpreter and the float64 tail, on the CPU -- against the float64 oracle driven by the
    restated Philox stream: same latent draws, same surviving rows, same x', log q, log w, same
    turn statistics (what tests/test_gpu_populate.py checks on the GPU)."""
    from oracle.flow_numpy import NumpyFlow
    from oracle.philox_numpy import latent_normals
    from oracle.populate_numpy import populate_turn

    g, cfg, sd = load_golden(name)
    sp = FlowSpec(cfg)
    theta = np.zeros(sp.n_theta, np.float32)
    ints = {}
    sp.load_state_dict_numpy(sd, theta, ints)
    prog = sp.fold(theta, ints).program(True)
    D, n, seed, offset = sp.D, 700, 0xABCDEF1234, 10**10 + 7
    rng = np.random.default_rng(3)
    scale, shift = rng.uniform(0.8, 1.6, D), rng.uniform(-0.3, 0.3, D)
    lo, hi = np.full(D, -3.5), np.full(D, 3.5)
    std = float(np.sqrt(sp.base_var))  # N(0, var I) base: z0 = std v, handed to the kernel as sqrt(T var)
    lpc, r_max = -D * np.log(7.0), 1.15 * np.sqrt(D) * sqrt_t * std
    ops = np.ascontiguousarray(prog.ops, dtype=np.int32)
    blob = np.ascontiguousarray(prog.blob, dtype=np.float32)
    xp = np.full((n, D), np.nan, dtype=np.float32)
    z = np.full((n, D), np.nan, dtype=np.float32)
    logq, logw = np.full(n, 7.0), np.full(n, 7.0)
    stats = np.array([-np.inf, 0.0])
    rc = simt_populate.simt_populate_draw(
        3, ops.ctypes.data, int(ops.shape[0]), blob.ctypes.data, D, sp.H, sp.activation, int(prog.final_buf),
        float(prog.const_logdet), n, seed, offset, r_max, sqrt_t * std, scale.ctypes.data, shift.ctypes.data, lo.ctypes.data,
        hi.ctypes.data, lpc, float("nan") if min_log_q is None else min_log_q, xp.ctypes.data, logq.ctypes.data,
        logw.ctypes.data, z.ctypes.data, stats.ctypes.data)
    assert rc == 0
    z_ref = latent_normals(seed, offset + np.arange(n), D) * std
    np.testing.assert_allclose(z, z_ref * sqrt_t, rtol=1e-5, atol=2e-5)  # the draw is the Philox stream
    kw = {k: v for k, v in dict(ftype=sp.ftype, net=sp.net, activation_name=cfg.get("activation", "relu"),
                                hidden_features=sp.H, base_var=sp.base_var).items()}
    if sp.ftype == "nsf":
        kw.update(num_bins=sp.num_bins, tail_bound=sp.tail_bound)
    nf = NumpyFlow(sd, **kw)
    t = populate_turn(nf, z_ref, scale=scale, shift=shift, lo=lo, hi=hi, log_prior_const=lpc, r_max=r_max,
                      sqrt_t=sqrt_t, min_log_q=min_log_q)
    # rows within fp32 rounding of the radius, a bound or min_log_q may flip
    zz = z_ref * sqrt_t
    edge = (np.abs(np.sqrt(np.sum(zz**2, axis=1)) - r_max) < 1e-4) | np.any(np.abs(np.abs(t["x"]) - 3.5) < 2e-3, axis=1)
    if min_log_q is not None:
        with np.errstate(invalid="ignore"):
            edge |= np.abs(np.nan_to_num(t["log_q"], nan=1e9) - min_log_q) < 1e-3
    dev_valid = ~np.isnan(logw)
    np.testing.assert_array_equal(dev_valid[~edge], t["valid"][~edge])
    assert 0.05 * n < t["valid"].sum() < 0.98 * n
    both = dev_valid & t["valid"]
    tol = dict(rtol=5e-4, atol=5e-4) if sp.ftype == "nsf" else dict(rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(xp[both].astype(np.float64) * scale + shift, t["x"][both], **tol)
    np.testing.assert_allclose(logq[both], t["log_q"][both], **tol)
    np.testing.assert_allclose(logw[both], t["log_w"][both], **tol)
    np.testing.assert_array_equal(np.isnan(logq), np.isnan(logw))
    assert stats[1] == dev_valid.sum() and stats[0] == logw[dev_valid].max()
