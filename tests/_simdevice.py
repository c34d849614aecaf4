"""TEST INFRASTRUCTURE: a simulated device for the HOST logic of ``nessai_b200.proposal``.

``SimLib`` stands in for ``libnessai_b200.so`` in CPU tests: each populate entry point of
``include/nessai_b200.h`` is answered by the float64 oracle (``oracle/``) reading and writing the
caller's buffers through the raw pointers it is given -- exactly the contract of the C ABI -- with
the tensors living in host memory.  It exists so that the Python side of the populate loops (slot
offsets, Philox counter bookkeeping, argument order, loop control, record layout, engine selection
in the nessai plugin) can be exercised without a GPU.  It is not a fallback: nothing under
``nessai_b200/`` refers to it, and the engines refuse to run without the real library unless a
test installs this one with ``install(monkeypatch)``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from oracle.philox_numpy import accept_uniform, latent_normals
from oracle.populate_numpy import populate_turn
from oracle.reparam_numpy import tail_rows


def _view(ptr, n, dtype):
    """Writable numpy view of ``n`` items at a raw address (``int`` / ``c_void_p`` / ``None``)."""
    if isinstance(ptr, C.c_void_p):
        ptr = ptr.value
    if not ptr:
        return None
    dtype = np.dtype(dtype)
    buf = (C.c_char * (int(n) * dtype.itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype, count=int(n))


class SimHandle:
    """What ``B200Flow._handle`` is to the real library: the flow the draw kernel evaluates."""

    def __init__(self, numpy_flow, D):
        self.flow, self.D = numpy_flow, int(D)
        self.value = id(self)


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass


class SimLib:
    def __init__(self):
        self.calls = []

    # ------------------------------------------------------------------ diagnostics
    def nb200_last_error(self):
        return b"simulated device"

    # ------------------------------------------------------------------ populate
    def nb200_populate_draw(self, handle, n, seed, row_offset, r_max, sqrt_t, sc, sh, lo, hi, lpc, min_log_q,
                            xp, logq, logw, z, stats, stream):
        self.calls.append(("draw", int(n), int(row_offset)))
        if n <= 0:
            return 0
        D = handle.D
        sqrt_t = float(sqrt_t) if sqrt_t > 0 else 1.0
        zz = latent_normals(int(seed), int(row_offset) + np.arange(n), D)
        mlq = None if (min_log_q is None or np.isnan(min_log_q) or min_log_q == -np.inf) else float(min_log_q)
        t = populate_turn(handle.flow, zz, scale=_view(sc, D, "f8"), shift=_view(sh, D, "f8"),
                          lo=_view(lo, D, "f8"), hi=_view(hi, D, "f8"),
                          log_prior_const=0.0 if np.isnan(lpc) else float(lpc), r_max=float(r_max), sqrt_t=sqrt_t,
                          min_log_q=mlq)
        with np.errstate(all="ignore"):
            xp_out, _ = handle.flow.inverse(zz * sqrt_t)
        _view(xp, n * D, "f4")[:] = xp_out.astype(np.float32).ravel()
        _view(logq, n, "f8")[:] = t["log_q"]
        _view(logw, n, "f8")[:] = t["log_w"]
        if z:
            _view(z, n * D, "f4")[:] = (zz * sqrt_t).astype(np.float32).ravel()
        st = _view(stats, 2, "f8")
        if t["valid"].any():
            st[0] = max(st[0], float(np.nanmax(t["log_w"])))
            st[1] += float(t["valid"].sum())
        return 0

    def nb200_reparam_tail(self, n, D, xp, kind, src, pa, pb, sc, sh, lo, hi, lpc, min_log_q, logq, logw, x64,
                           stats, stream):
        self.calls.append(("tail", int(n)))
        if n <= 0:
            return 0
        mlq = None if (np.isnan(min_log_q) or min_log_q == -np.inf) else float(min_log_q)
        lq = _view(logq, n, "f8")
        x, q, w, valid = tail_rows(_view(xp, n * D, "f4").reshape(n, D), lq.copy(), kind=_view(kind, D, "i4"),
                                   scale=_view(sc, D, "f8"), shift=_view(sh, D, "f8"), lo=_view(lo, D, "f8"),
                                   hi=_view(hi, D, "f8"), log_prior_const=0.0 if np.isnan(lpc) else float(lpc),
                                   min_log_q=mlq, pre_scale=_view(pa, D, "f8"), pre_shift=_view(pb, D, "f8"),
                                   src=_view(src, 3 * D, "i4"))
        _view(x64, n * D, "f8")[:] = x.ravel()
        lq[:] = q
        _view(logw, n, "f8")[:] = w
        st = _view(stats, 2, "f8")
        if valid.any():
            st[0] = max(st[0], float(w[valid].max()))
            st[1] += float(valid.sum())
        return 0

    def _accept(self, n, D, x, logw, logl, dmax, seed, row_offset, logp, tmpl, row_bytes, offs, logl_off, rows,
                cap, woff, counts):
        lw = _view(logw, n, "f8")
        mx = _view(dmax, 1, "f8")[0]
        u = accept_uniform(int(seed), int(row_offset) + np.arange(n))
        with np.errstate(all="ignore"):
            acc = ~np.isnan(lw) & ((lw - mx) > np.log(u))
        idx = np.flatnonzero(acc)
        m = int(min(len(idx), cap))
        offs = _view(offs, D + 1, "i4")
        rec = np.tile(_view(tmpl, row_bytes, "u1"), (m, 1))
        for d in range(D):
            rec[:, offs[d] : offs[d] + 8] = np.ascontiguousarray(x[idx[:m], d]).view(np.uint8).reshape(m, 8)
        if offs[D] >= 0:
            rec[:, offs[D] : offs[D] + 8] = np.full(m, float(logp)).view(np.uint8).reshape(m, 8)
        if logl and logl_off >= 0:
            ll = _view(logl, n, "f8")[idx[:m]]
            rec[:, logl_off : logl_off + 8] = np.ascontiguousarray(ll).view(np.uint8).reshape(m, 8)
        if m:
            _view(rows, (woff + m) * row_bytes, "u1")[woff * row_bytes :] = rec.ravel()
        c = _view(counts, 2, "i8")
        c[0], c[1] = len(idx), m
        return 0

    def nb200_populate_accept(self, n, D, xp, sc, sh, logw, logl, dmax, seed, row_offset, logp, tmpl, row_bytes,
                              offs, logl_off, rows, cap, woff, counts, scratch, stream):
        self.calls.append(("accept", int(n), int(row_offset)))
        if n <= 0:
            return 0
        x = _view(xp, n * D, "f4").reshape(n, D).astype(np.float64) * _view(sc, D, "f8") + _view(sh, D, "f8")
        return self._accept(n, D, x, logw, logl, dmax, seed, row_offset, logp, tmpl, row_bytes, offs, logl_off,
                            rows, cap, woff, counts)

    def nb200_populate_accept_x64(self, n, D, x64, logw, logl, dmax, seed, row_offset, logp, tmpl, row_bytes,
                                  offs, logl_off, rows, cap, woff, counts, scratch, stream):
        self.calls.append(("accept_x64", int(n), int(row_offset)))
        if n <= 0:
            return 0
        x = _view(x64, n * D, "f8").reshape(n, D)
        return self._accept(n, D, x, logw, logl, dmax, seed, row_offset, logp, tmpl, row_bytes, offs, logl_off,
                            rows, cap, woff, counts)

    def nb200_sum_exp(self, logw, n, dmax, partials, n_partials, stream):
        self.calls.append(("sum_exp", int(n)))
        p = _view(partials, n_partials, "f8")
        p[:] = 0.0
        if n > 0:
            lw = _view(logw, n, "f8")
            ok = ~np.isnan(lw) & (lw > -np.inf)
            p[0] = float(np.exp(lw[ok] - _view(dmax, 1, "f8")[0]).sum())
        return 0


def install(monkeypatch=None):
    """Route ``nessai_b200``'s library calls and CUDA stream queries to the simulation and make
    the engines' result hand-off host-only.  ``monkeypatch=None`` patches for the life of the
    process (spawned workers).  Returns the ``SimLib``."""
    import torch

    from nessai_b200 import _lib, proposal

    patch = monkeypatch.setattr if monkeypatch is not None else setattr
    sim = SimLib()
    patch(_lib, "load", lambda: sim)
    patch(torch.cuda, "current_stream", lambda device=None: _Stream())

    def gather_rows(self, n_local_written, n_samples):
        if self.world > 1:  # the all-gather of the accepted records, host tensors over gloo
            full, _ = proposal.gather_records(self.d_rows, int(n_local_written), int(n_samples), self.row_bytes,
                                              self.group)
            return full.numpy().copy().view(self.row_dtype)
        nb = int(n_local_written) * self.row_bytes
        return self.d_rows[:nb].numpy().copy().view(self.row_dtype)

    def run_serial_only(self, n_samples, drawsize, max_samples=1_000_000, host_prior=None, to_host=True):
        self._ensure(1, int(n_samples), False)
        return self._run_serial(int(n_samples), int(drawsize), max_samples, host_prior, False)

    patch(proposal.PopulateEngine, "_gather_rows", gather_rows)
    patch(proposal.PopulateEngine, "run", run_serial_only)  # the pipelined loop needs CUDA events
    return sim


def install_fake_cuda_async(monkeypatch=None):
    """Stand-ins for CUDA events / streams / pinned allocations so that the software-pipelined
    loop can run on the host: every "asynchronous" operation completes at once, which checks the
    loop's bookkeeping, not its stream ordering."""
    import contextlib

    import torch

    patch = monkeypatch.setattr if monkeypatch is not None else setattr

    class Event:
        def record(self, stream=None):
            pass

        def synchronize(self):
            pass

    class Stream:
        def __init__(self, device=None):
            pass

        def wait_event(self, ev):
            pass

        def synchronize(self):
            pass

    patch(torch.cuda, "Event", Event)
    patch(torch.cuda, "Stream", Stream)
    patch(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    for name in ("empty", "zeros"):
        orig = getattr(torch, name)
        patch(torch, name, lambda *a, _o=orig, **k: _o(*a, **{**k, "pin_memory": False}))


class SimFlowModel:
    """The two attributes ``PopulateEngine`` reads from a ``B200FlowModel``."""

    class _Model:
        def __init__(self, handle):
            import torch

            self.device = torch.device("cpu")
            self._handle = handle

        def _ready(self):
            pass

    def __init__(self, numpy_flow, D):
        self.model = self._Model(SimHandle(numpy_flow, D))
