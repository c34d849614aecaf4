"""The stand-alone coupling kernels (csrc/coupling.cuh: the HBM-roofline kernel of the bench) --
the CUDA source, unchanged -- run on the CPU under the SIMT shim of tests/_hostcheck and
compared with the numpy restatement the GPU test uses (tests/test_gpu_coupling.py)."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
from conftest import REPO, simt_or_skip
from test_gpu_coupling import numpy_coupling


@pytest.fixture(scope="module")
def simt_coupling(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d = os.path.join(REPO, "tests", "_hostcheck")
    out = tmp_path_factory.mktemp("simt") / "libcoupling_simt.so"
    res = subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", f"-I{d}/fake_cuda", "-o", str(out),
                          os.path.join(d, "coupling_kernels_simt.cpp")], capture_output=True, text=True)
    if res.returncode != 0:
        if "barrier" in res.stderr:
            pytest.skip("this g++ has no <barrier>")
        raise RuntimeError(res.stderr)
    lib = C.CDLL(str(out))
    lib.simt_coupling.restype = C.c_int
    lib.simt_coupling.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_void_p] + [C.c_int] * 4
    return simt_or_skip(lib, 256)


@pytest.mark.parametrize("D,tf,vec", [(16, list(range(1, 16, 2)), 1), (16, [3, 0, 9], 1), (8, [0, 2, 4, 6], 1),
                                      (4, [1, 3], 1), (32, list(range(0, 32, 2)), 1), (64, list(range(1, 64, 2)), 1),
                                      (16, list(range(0, 16, 2)), 0), (5, [4, 1], 0), (2, [1], 0)])
@pytest.mark.parametrize("additive", [0, 1])
@pytest.mark.parametrize("inverse", [0, 1])
def test_coupling_kernel_source_under_simt_shim(simt_coupling, D, tf, vec, additive, inverse):
    rng = np.random.default_rng(D + 7 * additive + 13 * inverse)
    n, grid = 333, 2  # ragged: the vector kernel's padded tail and its grid-stride loop both run
    d_tr = len(tf)
    x = rng.standard_normal((n, D)).astype(np.float32)
    params = rng.standard_normal((n, d_tr if additive else 2 * d_tr)).astype(np.float32)
    y = np.full((n, D), np.nan, dtype=np.float32)
    ld = np.full(n, np.nan, dtype=np.float32)
    tfa = np.asarray(tf, dtype=np.int32)
    rc = simt_coupling.simt_coupling(grid, x.ctypes.data, params.ctypes.data, y.ctypes.data, ld.ctypes.data, n, D,
                                     tfa.ctypes.data, d_tr, additive, inverse, vec)
    assert rc == 0
    y_ref, ld_ref = numpy_coupling(x, params, tf, bool(additive), bool(inverse))
    np.testing.assert_allclose(y, y_ref, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(ld, ld_ref, rtol=2e-5, atol=2e-5)
    ident = [f for f in range(D) if f not in tf]
    np.testing.assert_array_equal(y[:, ident], x[:, ident])  # identity features pass through bit for bit
