"""The rejection step + compaction kernel (csrc/accept.cuh: accept_fused_kernel -- the CUDA source
the product runs) on the CPU under the SIMT shim of tests/_hostcheck: accepted set, in-order
compaction across chunks (the look-back scan), capacity / write offset, record layout, the logL
field and the float64-row (x64) flavour, against numpy with the restated Philox uniforms."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
from conftest import REPO, simt_or_skip

from nessai_b200.livepoint import empty_structured_array, get_dtype
from oracle.philox_numpy import accept_uniform


@pytest.fixture(scope="module")
def simt_accept(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d = os.path.join(REPO, "tests", "_hostcheck")
    out = tmp_path_factory.mktemp("simt") / "libaccept_simt.so"
    res = subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", f"-I{d}/fake_cuda", "-o", str(out),
                          os.path.join(d, "accept_simt.cpp")], capture_output=True, text=True)
    if res.returncode != 0:
        if "barrier" in res.stderr:
            pytest.skip("this g++ has no <barrier>")
        raise RuntimeError(res.stderr)
    lib = C.CDLL(str(out))
    lib.simt_populate_accept.restype = C.c_int
    lib.simt_populate_accept.argtypes = ([C.c_int64, C.c_int] + [C.c_void_p] * 7 + [C.c_uint64, C.c_uint64, C.c_double,
                                         C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                         C.c_void_p, C.c_void_p])
    return simt_or_skip(lib, 256)


@pytest.mark.parametrize("n", [1, 1023, 1025, 5000])
@pytest.mark.parametrize("mode", ["affine", "x64", "affine_logl", "x64_logl"])
def test_accept_kernel_source_under_simt_shim(simt_accept, n, mode):
    D, seed, offset = 5, 0x5EED5EED5EED, 2**33 + 11
    rng = np.random.default_rng(n + len(mode))
    names = [f"p{i}" for i in range(D)]
    dtype = get_dtype(names)
    rb = dtype.itemsize
    logw = rng.normal(-2.0, 0.7, size=n)
    logw[rng.random(n) < 0.25] = np.nan
    if not np.any(~np.isnan(logw)):
        logw[0] = -1.0
    valid = ~np.isnan(logw)
    mx = np.array([logw[valid].max()])
    xp = rng.standard_normal((n, D)).astype(np.float32)
    scale, shift = rng.uniform(0.5, 2.0, D), rng.uniform(-1, 1, D)
    x64 = rng.standard_normal((n, D))
    logl = rng.normal(size=n)
    use_x64, use_logl = mode.startswith("x64"), mode.endswith("logl")
    x = x64 if use_x64 else xp.astype(np.float64) * scale + shift
    u = accept_uniform(seed, offset + np.arange(n))
    with np.errstate(invalid="ignore"):
        acc = valid & ((logw - mx[0]) > np.log(u))
    n_acc = int(acc.sum())
    tmpl = empty_structured_array(1, dtype=dtype).view(np.uint8).copy()
    offs = np.asarray([dtype.fields[nm][1] for nm in names] + [dtype.fields["logP"][1]], dtype=np.int32)
    for capacity, woff in ((n, 0), (max(n_acc // 2, 1), 3), (0, 0)):
        rows = np.full((woff + capacity + 2) * rb, 0xEE, dtype=np.uint8)
        counts = np.full(2, -1, dtype=np.int64)
        scratch = np.full(n // 1024 + 2, -1, dtype=np.int64)
        rc = simt_accept.simt_populate_accept(
            n, D, None if use_x64 else xp.ctypes.data, x64.ctypes.data if use_x64 else None, scale.ctypes.data,
            shift.ctypes.data, logw.ctypes.data, logl.ctypes.data if use_logl else None, mx.ctypes.data, seed, offset,
            -3.25, tmpl.ctypes.data, rb, offs.ctypes.data, int(dtype.fields["logL"][1]), rows.ctypes.data, capacity,
            woff, counts.ctypes.data, scratch.ctypes.data)
        assert rc == 0
        m = min(n_acc, capacity)
        assert counts.tolist() == [n_acc, m]
        rec = rows[woff * rb : (woff + m) * rb].view(dtype)
        got = np.stack([rec[nm] for nm in names], axis=-1) if m else np.empty((0, D))
        # draw order across chunk boundaries; the float64 rows are copied bit for bit, the affine ones are
        # formed by one fused multiply-add (the host build has none: equal to an ulp)
        if use_x64:
            np.testing.assert_array_equal(got, x[acc][:m])
        else:
            np.testing.assert_allclose(got, x[acc][:m], rtol=1e-15, atol=1e-15)
        assert np.all(rec["logP"] == -3.25) and np.all(rec["it"] == 0)
        if use_logl:
            np.testing.assert_array_equal(rec["logL"], logl[acc][:m])
        else:
            assert np.all(np.isnan(rec["logL"]))
        # nothing outside [write_offset, write_offset + m) is touched
        assert np.all(rows[: woff * rb] == 0xEE) and np.all(rows[(woff + m) * rb :] == 0xEE)
