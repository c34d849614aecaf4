"""TEST INFRASTRUCTURE: the simulated device of tests/_simdevice.py with the populate entry points
answered by the product's OWN CUDA sources, compiled for the CPU against the SIMT shim of
tests/_hostcheck (generic draw kernel, non-affine tail, sum-exp, rejection + compaction) instead
of by the oracle.  Together with the real Python engines this is the whole populate path minus
nvcc and the tcgen05 specialisations, runnable without a GPU -- and checked AGAINST the oracle,
never used in its place."""

from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

import _simdevice

HERE = os.path.dirname(os.path.abspath(__file__))
HOSTCHECK = os.path.join(HERE, "_hostcheck")


def build(tmpdir):
    """Compile the four harnesses; returns the loaded libraries (or None without a C++20 g++)."""
    gxx = shutil.which("g++")
    if gxx is None:
        return None
    libs = {}
    for name, srcs in (("populate", ["populate_draw_simt.cpp"]), ("accept", ["accept_simt.cpp"]),
                       ("tail", ["reparam_kernels_simt.cpp", "reparam_host.cpp"])):
        out = os.path.join(str(tmpdir), f"lib{name}_simt.so")
        res = subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", f"-I{HOSTCHECK}/fake_cuda", "-o",
                              out] + [os.path.join(HOSTCHECK, s) for s in srcs], capture_output=True, text=True)
        if res.returncode != 0:
            if "barrier" in res.stderr:
                return None
            raise RuntimeError(res.stderr)
        libs[name] = C.CDLL(out)
    vp, i32, i64, u64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double
    libs["populate"].simt_populate_draw.argtypes = ([i32, vp, i32, vp, i32, i32, i32, i32, f64, i64, u64, u64, f32, f32]
                                                    + [vp] * 4 + [f64, f64] + [vp] * 5)
    libs["accept"].simt_populate_accept.argtypes = ([i64, i32] + [vp] * 7 + [u64, u64, f64, vp, i32, vp, i32, vp, i64,
                                                    i64, vp, vp])
    libs["tail"].simt_reparam_tail.argtypes = [i32, i64, i32] + [vp] * 9 + [f64, f64] + [vp] * 4
    libs["tail"].simt_sum_exp.argtypes = [i32, vp, i64, vp, vp]
    for lib, fn in (("populate", "simt_populate_draw"), ("accept", "simt_populate_accept")):
        getattr(libs[lib], fn).restype = i32
    libs["tail"].simt_reparam_tail.restype = None
    libs["tail"].simt_sum_exp.restype = None
    if any(lib.simt_probe(256) != 0 for lib in libs.values()):
        return None  # this machine cannot hold a block of OS threads
    return libs


class SimtHandle:
    """The folded inverse program of a flow, as ``nb200_flow_set_program`` receives it."""

    def __init__(self, spec, prog):
        self.D, self.H, self.activation = spec.D, spec.H, spec.activation
        self.ops = np.ascontiguousarray(prog.ops, dtype=np.int32)
        self.blob = np.ascontiguousarray(prog.blob, dtype=np.float32)
        self.final_buf, self.const_logdet = int(prog.final_buf), float(prog.const_logdet)
        self.value = id(self)


def _p(ptr):
    if isinstance(ptr, C.c_void_p):
        return ptr.value
    return None if not ptr else int(ptr)


class SimtLib(_simdevice.SimLib):
    def __init__(self, libs, grid=3):
        super().__init__()
        self.libs, self.grid = libs, grid

    def nb200_populate_draw(self, handle, n, seed, row_offset, r_max, sqrt_t, sc, sh, lo, hi, lpc, min_log_q,
                            xp, logq, logw, z, stats, stream):
        self.calls.append(("draw", int(n), int(row_offset)))
        if n <= 0:
            return 0
        h = handle
        return self.libs["populate"].simt_populate_draw(
            self.grid, h.ops.ctypes.data, int(h.ops.shape[0]), h.blob.ctypes.data, h.D, h.H, h.activation, h.final_buf,
            h.const_logdet, int(n), int(seed), int(row_offset), float(r_max), float(sqrt_t), _p(sc), _p(sh), _p(lo),
            _p(hi), float(lpc), float(min_log_q), _p(xp), _p(logq), _p(logw), _p(z), _p(stats))

    def nb200_reparam_tail(self, n, D, xp, kind, src, pa, pb, sc, sh, lo, hi, lpc, min_log_q, logq, logw, x64,
                           stats, stream):
        self.calls.append(("tail", int(n)))
        if n <= 0:
            return 0
        self.libs["tail"].simt_reparam_tail(
            self.grid, int(n), int(D), _p(xp), _p(kind), _p(src), _p(pa), _p(pb), _p(sc), _p(sh), _p(lo), _p(hi),
            0.0 if np.isnan(lpc) else float(lpc), -np.inf if np.isnan(min_log_q) else float(min_log_q), _p(logq),
            _p(logw), _p(x64), _p(stats))
        return 0

    def _accept_simt(self, tag, n, D, xp, x64, sc, sh, logw, logl, dmax, seed, row_offset, logp, tmpl, row_bytes, offs,
                     logl_off, rows, cap, woff, counts, scratch):
        self.calls.append((tag, int(n), int(row_offset)))
        if n <= 0:
            return 0
        return self.libs["accept"].simt_populate_accept(
            int(n), int(D), _p(xp), _p(x64), _p(sc), _p(sh), _p(logw), _p(logl), _p(dmax), int(seed), int(row_offset),
            float(logp), _p(tmpl), int(row_bytes), _p(offs), int(logl_off), _p(rows), int(cap), int(woff), _p(counts),
            _p(scratch))

    def nb200_populate_accept(self, n, D, xp, sc, sh, logw, logl, dmax, seed, row_offset, logp, tmpl, row_bytes,
                              offs, logl_off, rows, cap, woff, counts, scratch, stream):
        return self._accept_simt("accept", n, D, xp, None, sc, sh, logw, logl, dmax, seed, row_offset, logp, tmpl,
                                 row_bytes, offs, logl_off, rows, cap, woff, counts, scratch)

    def nb200_populate_accept_x64(self, n, D, x64, logw, logl, dmax, seed, row_offset, logp, tmpl, row_bytes,
                                  offs, logl_off, rows, cap, woff, counts, scratch, stream):
        return self._accept_simt("accept_x64", n, D, None, x64, x64, x64, logw, logl, dmax, seed, row_offset, logp, tmpl,
                                 row_bytes, offs, logl_off, rows, cap, woff, counts, scratch)

    def nb200_sum_exp(self, logw, n, dmax, partials, n_partials, stream):
        self.calls.append(("sum_exp", int(n)))
        self.libs["tail"].simt_sum_exp(int(n_partials), _p(logw), int(max(n, 0)), _p(dmax), _p(partials))
        return 0


def install(monkeypatch, libs):
    """As ``_simdevice.install`` with the CUDA-source-backed library."""
    from nessai_b200 import _lib

    _simdevice.install(monkeypatch)
    sim = SimtLib(libs)
    monkeypatch.setattr(_lib, "load", lambda: sim)
    return sim


class SimtFlowModel:
    """What ``PopulateEngine`` reads from a ``B200FlowModel``: a device and the flow handle."""

    def __init__(self, spec, prog):
        import torch

        handle = SimtHandle(spec, prog)

        class _Model:
            device = torch.device("cpu")
            _handle = handle

            @staticmethod
            def _ready():
                pass

        self.model = _Model()
