"""Device-resident ``populate()`` and the standalone ``B200FlowProposal``.

``PopulateEngine`` is the B200 version of the loop body of
``FlowProposal.populate``
(/root/reference/src/nessai/proposal/flowproposal/flowproposal.py:431-510): each
turn is ONE fused kernel (Philox latent draw -> latent-radius truncation ->
inverse flow -> diagonal inverse rescale -> prior bounds -> log-weights -> max)
followed by the rejection step + in-order compaction into structured live-point
records on the device; only the accepted records cross PCIe, already in the
sampler's dtype.  With ``torch.distributed`` initialised (one process per GPU)
the draw is sharded over ranks by global row index; the ranks exchange the
scalar max / counts and all-gather the accepted records over NCCL.

``B200FlowProposal`` mirrors the part of ``BaseFlowProposal`` / ``FlowProposal``
on the hot path (same method names, arguments and attributes; SURVEY.md 8a
b1-b11) without importing nessai, for the default z-score / null
reparameterisation.  ``nessai_b200.nessai_plugin`` binds the same engine into the
real ``nessai.proposal.FlowProposal`` when nessai is installed.
"""

from __future__ import annotations

import ctypes as C
import datetime
import logging
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from .flowmodel import B200FlowModel, _ptr, _stream
from .livepoint import (
    NON_SAMPLING_DEFAULTS,
    NON_SAMPLING_PARAMETERS,
    empty_structured_array,
    get_dtype,
    live_points_to_array,
)

logger = logging.getLogger(__name__)


def compute_radius(n, q=0.95):
    """/root/reference/src/nessai/utils/sampling.py:15-33."""
    from scipy import stats

    return stats.chi.ppf(q, n)


def _dist_info(group=None):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def detect_uniform_box_prior(model, rng, n=256, atol=1e-9):
    """Return ``log p`` if the model's prior is uniform on its (finite) box, else
    ``None``.  Checked numerically on random in-bounds points against
    ``-sum(log(hi - lo))``; a model can also declare it with an attribute
    ``uniform_box_prior = True`` / ``False``."""
    declared = getattr(model, "uniform_box_prior", None)
    names = list(model.names)
    lo = np.array([model.bounds[n_][0] for n_ in names], dtype=np.float64)
    hi = np.array([model.bounds[n_][1] for n_ in names], dtype=np.float64)
    if not (np.all(np.isfinite(lo)) and np.all(np.isfinite(hi))):
        return None
    const = -float(np.sum(np.log(hi - lo)))
    if declared is False:
        return None
    if declared is True:
        return const
    x = empty_structured_array(n, names)
    u = rng.random((n, len(names)))
    for i, nm in enumerate(names):
        x[nm] = lo[i] + (hi[i] - lo[i]) * u[:, i]
    try:
        lp = np.asarray(model.log_prior(x), dtype=np.float64)
    except Exception:
        return None
    if lp.shape != (n,) or not np.all(np.abs(lp - const) <= atol * max(1.0, abs(const))):
        return None
    return const


def shard_rows(n_total: int, rank: int, world: int):
    """Contiguous shard of a ``n_total``-row draw: ``(n_local, first_row)``."""
    base, rem = divmod(int(n_total), world)
    return base + (1 if rank < rem else 0), rank * base + min(rank, rem)


def gather_records(local_rows: torch.Tensor, n_local: int, n_keep: int, row_bytes: int, group=None,
                   to_host: bool = False):
    """All-gather the accepted live-point records (uint8, ``row_bytes`` each) of every
    rank, rank-major, and keep the first ``n_keep``.  Works on any backend (NCCL on
    GPUs; gloo in the CPU tests).  Returns ``(rows, counts)``: a uint8 tensor on
    ``local_rows.device`` -- or, with ``to_host``, a uint8 tensor in PINNED host memory
    filled by one asynchronous device-to-host copy per rank segment (no device-side
    concatenation, no pageable staging)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = local_rows.device
    cnt = torch.tensor([int(n_local)], dtype=torch.int64, device=dev)
    if dev.type == "cuda":
        allc_t = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc_t, cnt, group=group)
        allc = [int(c) for c in allc_t.cpu().tolist()]  # one synchronisation
    else:
        tmp = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(tmp, cnt, group=group)
        allc = [int(c.item()) for c in tmp]
    mx = max(max(allc), 1)
    if local_rows.numel() >= mx * row_bytes and dev.type == "cuda":
        send = local_rows[: mx * row_bytes]  # the tail past n_local is never read back
    else:
        send = torch.zeros(mx * row_bytes, dtype=torch.uint8, device=dev)
        send[: n_local * row_bytes] = local_rows[: n_local * row_bytes]
    if dev.type == "cuda":
        recv_t = torch.empty((world, mx * row_bytes), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(recv_t, send, group=group)
        recv = [recv_t[r] for r in range(world)]
    else:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send, group=group)
    # rank-major, first n_keep records
    take, left = [], int(n_keep)
    for r in range(world):
        k = min(allc[r], left)
        take.append(k)
        left -= k
    total = sum(take)
    if to_host and dev.type == "cuda":
        timing = bool(os.environ.get("NB200_TIMING"))
        if timing:
            import sys
            import time

            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
        host = torch.empty(total * row_bytes, dtype=torch.uint8, pin_memory=True)
        if timing and dist.get_rank(group) == 0:
            print(f"[nb200 timing] pinned alloc {1e3 * (time.perf_counter() - t0):.2f} ms for "
                  f"{total * row_bytes / 1e6:.1f} MB", file=sys.stderr)
        off = 0
        for r in range(world):
            nb = take[r] * row_bytes
            if nb:
                host[off : off + nb].copy_(recv[r][:nb], non_blocking=True)
                off += nb
        torch.cuda.current_stream(dev).synchronize()
        return host, allc
    parts = [recv[r][: take[r] * row_bytes] for r in range(world)]
    return torch.cat(parts), allc


class IndexPool:
    """List-like pool of sample indices (``pop`` / ``len`` / truthiness / iteration) over a
    numpy permutation.  The reference keeps ``rng.permutation(n).tolist()``
    (/root/reference/src/nessai/proposal/flowproposal/flowproposal.py:530) and pops from
    its end; for a 1e6-row pool building and freeing that list of Python ints costs
    milliseconds per populate -- more than the fused draw itself -- so the same order is
    served from the array."""

    __slots__ = ("_a", "_n")

    def __init__(self, permutation):
        self._a = np.asarray(permutation)
        self._n = int(self._a.size)

    def pop(self):
        if self._n == 0:
            raise IndexError("pop from empty pool")
        self._n -= 1
        return int(self._a[self._n])

    def __len__(self):
        return self._n

    def __bool__(self):
        return self._n > 0

    def __iter__(self):
        return iter(self._a[: self._n].tolist())

    def tolist(self):
        return self._a[: self._n].tolist()


class AccumulateControl:
    """Loop control of ``populate`` with ``accumulate_weights=True``, host side, exactly the
    reference's (/root/reference/src/nessai/proposal/flowproposal/flowproposal.py:431-490,
    504-512): a turn that added no surviving row only checks ``max_samples`` (``continue``,
    :436-439,449-453); otherwise the rejection step runs when the expected pool size has reached
    ``n_samples`` (:477-485); ``max_samples`` ends the loop after the turn (:486-488); after the
    loop the rejection step is repeated if rows were added since the last one (:505-507)."""

    def __init__(self, n_samples: int, max_samples: int):
        import math

        self.n_samples, self.max_samples = int(n_samples), int(max_samples)
        self.log_n = math.log(self.n_samples) if self.n_samples > 0 else -math.inf
        self.n_proposed = self.n_accepted = 0
        self.stale = False  # rows were added since the last rejection step
        self.stopped_on_max_samples = False

    def go_on(self) -> bool:
        return not self.stopped_on_max_samples and self.n_accepted < self.n_samples

    def turn_drawn(self, n_drawn: int, added_rows: bool, n_expected: float) -> bool:
        """Account for a turn; True when the rejection step must run now."""
        import math

        self.n_proposed += int(n_drawn)
        if not added_rows:
            return False
        self.stale = True
        return n_expected > 0.0 and math.log(n_expected) >= self.log_n

    def rejected(self, n_accepted: int):
        self.n_accepted = int(n_accepted)
        self.stale = False

    def end_turn(self):
        if self.n_proposed > self.max_samples:
            self.stopped_on_max_samples = True


class PopulateEngine:
    """Fused populate turns on one GPU (optionally one rank of many)."""

    def __init__(self, flow: B200FlowModel, names, row_dtype: np.dtype, group=None, row_template=None):
        self.flow = flow
        self.model = flow.model
        self.device = self.model.device
        self.names = list(names)
        self.D = len(self.names)
        self.row_dtype = np.dtype(row_dtype)
        self.row_bytes = self.row_dtype.itemsize
        if self.row_bytes % 4:
            raise NotImplementedError(
                f"live-point dtype itemsize {self.row_bytes} is not a multiple of 4"
            )
        offs = [self.row_dtype.fields[n][1] for n in self.names]
        for n in self.names:
            if self.row_dtype.fields[n][0] != np.dtype("f8"):
                raise NotImplementedError("live-point parameters must be float64")
        logp_off = self.row_dtype.fields["logP"][1] if "logP" in self.row_dtype.names else -1
        self.field_offsets = np.asarray(offs + [logp_off], dtype=np.int32)
        self.logl_offset = (
            int(self.row_dtype.fields["logL"][1])
            if "logL" in self.row_dtype.names and self.row_dtype.fields["logL"][0] == np.dtype("f8")
            else -1
        )
        tmpl = row_template if row_template is not None else empty_structured_array(1, dtype=self.row_dtype)
        tmpl = np.ascontiguousarray(tmpl, dtype=self.row_dtype).reshape(1)
        self.d_template = torch.from_numpy(tmpl.view(np.uint8).copy()).to(self.device)
        self.group = group
        self.rank, self.world = _dist_info(group)
        self._cap = 0
        self._rows_cap = 0
        self._gen = 0  # bumped whenever a device buffer is (re)allocated
        self._turn_rows = 0  # global rows drawn so far (Philox counter base)
        # the first n_model parameters are the model's (what a device likelihood is given); the rest
        # are auxiliary x-space parameters of the reparameterisation (flowproposal/base.py:539-546)
        self.n_model = self.D
        self.seed = None
        self.min_log_q, self.likelihood, self.log_l_threshold = -float("inf"), None, -float("inf")

    # ---------------------------------------------------------------- buffers
    def _ensure(self, n_local: int, capacity: int, want_z: bool):
        dev = self.device
        if n_local > self._cap:
            self.d_xp = torch.empty((n_local, self.D), dtype=torch.float32, device=dev)
            self.d_logq = torch.empty(n_local, dtype=torch.float64, device=dev)
            self.d_logw = torch.empty(n_local, dtype=torch.float64, device=dev)
            self.d_scratch = torch.empty(n_local // 1024 + 2, dtype=torch.int64, device=dev)
            self.d_z = None
            self._cap = n_local
            self._gen += 1
        if want_z and (self.d_z is None or self.d_z.shape[0] < n_local):
            self.d_z = torch.empty((self._cap, self.D), dtype=torch.float32, device=dev)
            self._gen += 1
        if capacity > self._rows_cap:
            self.d_rows = torch.empty(capacity * self.row_bytes, dtype=torch.uint8, device=dev)
            self._rows_cap = capacity
            self._gen += 1
        if not hasattr(self, "d_stats"):
            self.d_stats = torch.empty(2, dtype=torch.float64, device=dev)
            self._stats_init = torch.tensor([-float("inf"), 0.0], dtype=torch.float64, device=dev)
            self.d_counts = torch.zeros(2, dtype=torch.int64, device=dev)
            self._gen += 1

    def configure(self, scale, shift, lo, hi, log_prior_const, r_max, sqrt_temperature=1.0, min_log_q=None,
                  likelihood=None, log_l_threshold=None):
        """``min_log_q``: MinLogQTruncation (truncation.py:368-394).  ``likelihood``: a callable
        ``x (n, D) float64 device tensor -> (n,) log-likelihood`` evaluated on every proposed
        row inside the loop; with ``log_l_threshold`` rows at or below it are dropped before the
        rejection step (LikelihoodThresholdTruncation, truncation.py:397-429) and the accepted
        records carry their logL."""
        # one H2D copy of the four float64 vectors, and only when they change (populate() calls
        # this every time; the z-score statistics only change when the flow is retrained)
        c = getattr(self, "_cfg_host", None)
        same = c is not None and all(
            a is b or (np.shape(a) == b.shape and np.array_equal(a, b)) for a, b in zip((scale, shift, lo, hi), c)
        )
        if not same:
            host = [np.array(a, dtype=np.float64) for a in (scale, shift, lo, hi)]
            dev = torch.from_numpy(np.stack(host)).to(self.device)
            self.d_scale, self.d_shift, self.d_lo, self.d_hi = dev[0], dev[1], dev[2], dev[3]
            self._cfg_host = host
            self._gen += 1
        self.log_prior_const = log_prior_const
        self.r_max = float(r_max) if r_max else 0.0
        self.sqrt_t = float(sqrt_temperature)
        self.min_log_q = -float("inf") if min_log_q is None or np.isnan(min_log_q) else float(min_log_q)
        self.likelihood = likelihood
        self.log_l_threshold = -float("inf") if log_l_threshold is None else float(log_l_threshold)

    def set_seed(self, seed: int) -> None:
        """Restart the Philox streams: ``seed`` keys them, the row counter starts at zero (the
        cached argument lists of the draw / accept calls carry the seed, so they are dropped)."""
        self.seed = int(seed)
        self._turn_rows = 0
        self._draw_key = self._accept_key = None

    def _seed(self):
        if self.seed is None:
            s = torch.randint(0, 2**62, (1,), dtype=torch.int64)
            if self.world > 1:
                import torch.distributed as dist

                s = s.to(self.device)
                dist.broadcast(s, src=0, group=self.group)
                s = s.cpu()
            self.seed = int(s.item())
        return self.seed

    def _shard(self, n_total: int):
        return shard_rows(n_total, self.rank, self.world)

    # ------------------------------------------------------------------ turns
    def _call(self, fn, args, what):
        idx = self.device.index
        if idx is None or torch.cuda.current_device() == idx:
            rc = fn(*args)
        else:
            with torch.cuda.device(self.device):
                rc = fn(*args)
        if rc:
            _lib.check(rc, what)

    def draw_turn(self, n_total: int, want_z: bool = False):
        """One fused draw of ``n_total`` global rows (this rank's shard).
        Leaves x / log_q / log_w on the device; returns ``n_local``."""
        self.model._ready()
        key = (n_total, want_z, self._gen, self.model._handle.value, self.log_prior_const, self.r_max, self.sqrt_t,
               self.min_log_q)
        if getattr(self, "_draw_key", None) != key:
            # the argument list only changes with the buffers / configuration: build it once
            n_local, start = self._shard(n_total)
            self._ensure(max(n_local, 1), self._rows_cap, want_z)
            key = key[:2] + (self._gen,) + key[3:]
            lpc = float("nan") if self.log_prior_const is None else float(self.log_prior_const)
            self._draw_args = [
                self.model._handle, n_local, self._seed(), 0, self.r_max, self.sqrt_t,
                self.d_scale.data_ptr(), self.d_shift.data_ptr(), self.d_lo.data_ptr(), self.d_hi.data_ptr(),
                lpc, self.min_log_q, self.d_xp.data_ptr(), self.d_logq.data_ptr(), self.d_logw.data_ptr(),
                self.d_z.data_ptr() if want_z else None, self.d_stats.data_ptr(), None,
            ]
            self._draw_shard = (n_local, start)
            self._draw_key = key
            self._draw_fn = _lib.load().nb200_populate_draw
        n_local, start = self._draw_shard
        self.d_stats.copy_(self._stats_init, non_blocking=True)  # {max log_w = -inf, n_valid = 0}
        args = self._draw_args
        args[3] = self._turn_rows + start
        args[17] = torch.cuda.current_stream(self.device).cuda_stream
        self._call(self._draw_fn, args, "nb200_populate_draw")
        self._last = self._draw_shard
        self._after_draw(n_local)
        return n_local

    def _after_draw(self, n_local: int):
        """What follows the fused draw kernel within a turn (overridden by GeneralPopulateEngine)."""
        if self.likelihood is not None:
            self._apply_device_likelihood(n_local)

    def device_log_likelihood(self, n_written: int, fn) -> torch.Tensor:
        """``fn(x)`` on the accepted records still resident in ``d_rows`` (``x``: ``(n, D)``
        float64, gathered from the records' parameter fields)."""
        words = self.row_bytes // 4
        rows32 = self.d_rows[: n_written * self.row_bytes].view(torch.int32).view(n_written, words)
        if getattr(self, "_param_cols", None) is None or self._param_cols.numel() != 2 * self.n_model:
            offs = self.field_offsets[: self.n_model] // 4
            cols = np.stack([offs, offs + 1], axis=1)
            self._param_cols = torch.from_numpy(cols.reshape(-1).astype(np.int64)).to(self.device)
        x = rows32.index_select(1, self._param_cols).contiguous().view(torch.float64)
        return torch.as_tensor(fn(x), device=self.device).to(torch.float64).reshape(n_written)

    def sharded_pool_likelihood(self, rows: np.ndarray, fn) -> bool:
        """logL of a pool that the ranks of a node assembled in shared host memory: every rank
        evaluates ``fn`` on the records IT accepted -- still resident in its HBM -- and writes their
        ``logL`` fields in place, instead of every rank evaluating the whole pool on its host
        (flowproposal.py:519-523).  Collective; False when the last populate did not use the shared pool."""
        shared = getattr(self, "_last_shared", None)
        if self.world == 1 or shared is None:
            return False
        n_local = sum(hi - lo for lo, hi, _ in self._segments)
        if n_local:
            first = min(src for _, _, src in self._segments)
            last = max(src + hi - lo for lo, hi, src in self._segments)
            ll = self.device_log_likelihood(last, fn)
            host = torch.empty(last, dtype=torch.float64, pin_memory=True)
            host.copy_(ll, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            vals = host.numpy()
            assert first >= 0
            col = rows["logL"]
            for lo, hi, src in self._segments:
                col[lo:hi] = vals[src : src + hi - lo]
        shared.barrier()  # every rank's values have landed
        return True

    def _apply_device_likelihood(self, n: int):
        """logL of every proposed row on the device (flowproposal.py:456-467 without the host
        round trip); rows at or below the threshold leave the turn, and the statistics of the
        rejection step (max log_w, valid count) are recomputed over what is left."""
        if n <= 0:
            return
        self.likelihood_evaluations = getattr(self, "likelihood_evaluations", 0) + n
        if getattr(self, "d_logl", None) is None or self.d_logl.shape[0] < self._cap:
            self.d_logl = torch.empty(self._cap, dtype=torch.float64, device=self.device)
            self._gen += 1
        self._likelihood_cut(self.physical_x(n), 0, n, n)

    def _likelihood_cut(self, x: torch.Tensor, off: int, n: int, n_stats: int):
        """logL of the ``n`` rows at row offset ``off`` (physical ``x``), the contour cut on their
        weights, and the statistics recomputed over the first ``n_stats`` rows of the weights."""
        if self.n_model < self.D:
            x = x[:, : self.n_model].contiguous()
        ll = torch.as_tensor(self.likelihood(x), device=self.device).to(torch.float64).reshape(n)
        self.d_logl[off : off + n] = ll
        self.d_logw[off : off + n].masked_fill_(~(ll > self.log_l_threshold), float("nan"))
        lw = self.d_logw[:n_stats]
        valid = ~torch.isnan(lw)
        self.d_stats[0] = torch.where(valid, lw, torch.full_like(lw, -float("inf"))).max()
        self.d_stats[1] = valid.sum().to(torch.float64)

    def physical_x(self, n: int) -> torch.Tensor:
        """x = x' * scale + shift (float64) of the last draw, on the device."""
        return self.d_xp[:n].to(torch.float64) * self.d_scale + self.d_shift

    def accept_turn(self, capacity_left: int, write_offset: int):
        """Rejection + compaction of the last draw.  Returns the number of rows
        accepted on this rank (device scalar tensor ``d_counts``)."""
        n_local, start = self._last
        if self.world > 1:
            # the rejection step normalises by the maximum over the WHOLE turn
            x = self._peer_exchange()
            if x is not None:
                x.allgather(x.KIND_MAX, self.d_stats, 1, None, self.d_stats)
            else:
                import torch.distributed as dist

                dist.all_reduce(self.d_stats[0:1], op=dist.ReduceOp.MAX, group=self.group)
        with_logl = self.likelihood is not None and getattr(self, "d_logl", None) is not None and self.logl_offset >= 0
        key = (n_local, self._gen, self.log_prior_const, with_logl)
        if getattr(self, "_accept_key", None) != key:
            lp = 0.0 if self.log_prior_const is None else float(self.log_prior_const)
            self._accept_args = [
                n_local, self.D, self.d_xp.data_ptr(), self.d_scale.data_ptr(), self.d_shift.data_ptr(),
                self.d_logw.data_ptr(), self.d_logl.data_ptr() if with_logl else None,
                self.d_stats.data_ptr(), self._seed(), 0, lp,
                self.d_template.data_ptr(), self.row_bytes, self.field_offsets.ctypes.data, self.logl_offset,
                self.d_rows.data_ptr(), 0, 0, self.d_counts.data_ptr(), self.d_scratch.data_ptr(), None,
            ]
            self._accept_key = key
            self._accept_fn = _lib.load().nb200_populate_accept
        args = self._accept_args
        args[9] = self._turn_rows + start
        args[16] = int(capacity_left)
        args[17] = int(write_offset)
        args[20] = torch.cuda.current_stream(self.device).cuda_stream
        self._call(self._accept_fn, args, "nb200_populate_accept")
        return self.d_counts

    def _peer_exchange(self):
        """The peer-memory exchange of this engine's ranks (xchg.py), set up collectively on first
        use; None when it is unavailable (another node, no CUDA IPC, NB200_NO_XCHG=1, CPU tests)."""
        if not hasattr(self, "_xchg"):
            from .xchg import make_peer_exchange

            self._xchg = make_peer_exchange(self.device, self.group) if self.device.type == "cuda" else None
        return self._xchg

    def run(self, n_samples: int, drawsize: int, max_samples: int = 1_000_000, host_prior=None,
            to_host: bool = True):
        """The whole ``while n_accepted < n_samples`` loop.

        Returns ``(rows, n_proposed, n_accepted)``; ``rows`` is a fresh
        structured numpy array (at most ``n_samples`` records, draw order,
        rank-major when sharded).  ``host_prior(x_struct) -> log_p`` is used
        when the prior is not a uniform box (evaluated on the host like the
        reference does, /root/reference/.../flowproposal/base.py:1032-1051).

        With the prior on the device the loop is software-pipelined: the counts of
        turn t are read through a pinned buffer and an event, and while the host waits
        for them the draw of turn t + 1 is already running when the loop is expected
        to go on (it goes on exactly as the reference's does -- a draw that turns out
        not to be needed is discarded and the Philox counter is not advanced for it, so
        the pool does not depend on the speculation).  On one GPU the accepted records
        of turn t also cross to the host on a copy stream under the next draw.
        ``to_host=False`` leaves the records in ``d_rows`` (``rows`` is then the number of
        records written on this rank): the device side of the loop alone.
        """
        n_samples, drawsize = int(n_samples), int(drawsize)
        self._ensure(1, n_samples, False)
        timing = bool(os.environ.get("NB200_TIMING"))
        if host_prior is not None or (timing and to_host) or n_samples <= 0:
            return self._run_serial(n_samples, drawsize, max_samples, host_prior, timing)
        dev = self.device
        if not hasattr(self, "_h_counts"):
            self._h_counts = torch.zeros(2 * max(self.world, 2), dtype=torch.int64, pin_memory=True)
            self._h_counts_np = self._h_counts.numpy()
            self._ev_counts = torch.cuda.Event()
            self._copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        tr = [] if os.environ.get("NB200_TRACE") else None  # host timeline of this call
        if tr is not None:
            import time

            self.last_trace = tr
            tr.append(("start", time.perf_counter()))
        n_proposed = n_accepted = n_local_written = 0
        hint = getattr(self, "_accept_hint", None)  # accepted rows per turn, last seen
        drawn_ahead = False
        rb = self.row_bytes
        max_turns = int(max_samples) // drawsize + 1
        # (a small pool -- the reference's default poolsize is nlive -- is copied once at the end: a
        # pinned block and a copy-stream hop per turn cost more than its few hundred KB are worth)
        single = self.world == 1 and to_host and n_samples * rb > (2 << 20)
        cs = self._copy_stream
        # several GPUs on one node: the pool is assembled in host memory shared by the ranks,
        # each rank copying only its own records (hostpool.py); else all-gather at the end
        shared = self._shared_pool(n_samples * rb) if (self.world > 1 and to_host) else None
        if shared is not None:
            pool_host = shared.tensors[self._pool_turn % shared.n_buffers]
            self._pool_turn += 1
        placed = 0  # records of all ranks placed in the shared pool so far
        self._segments = []  # (pool row lo, hi, first local record) of this rank's share, per turn
        self._last_shared = shared
        # pinned destination of the accepted records (one GPU): sized for what the turns are
        # expected to add; if that turns out too small the records are copied once at the end
        host, host_cap, overflow, copied = None, 0, False, 0
        if single and hint:
            host_cap = min(n_samples, int(1.3 * hint * max_turns) + 4096)
            host = torch.empty(host_cap * rb, dtype=torch.uint8, pin_memory=True)

        def copy_rows(lo, hi):
            with torch.cuda.stream(cs):
                host[lo * rb : hi * rb].copy_(self.d_rows[lo * rb : hi * rb], non_blocking=True)

        turn = 0
        while True:
            if not drawn_ahead:
                self.draw_turn(drawsize)
            drawn_ahead = False
            n_proposed += drawsize
            turn += 1
            counts = self.accept_turn(n_samples - n_local_written, n_local_written)
            self._turn_rows += drawsize
            if self.world == 1:
                self._h_counts[:2].copy_(counts, non_blocking=True)
            else:
                import torch.distributed as dist

                # every rank's {accepted, written}: the global count and this rank's offset
                if getattr(self, "_d_allc", None) is None:
                    self._d_allc = torch.empty(2 * self.world, dtype=torch.int64, device=dev)
                x = self._peer_exchange()
                if x is not None:
                    x.allgather(x.KIND_COUNTS, counts, 2, self._d_allc, None)
                    x.stage_error_flag()
                else:
                    dist.all_gather_into_tensor(self._d_allc, counts, group=self.group)
                self._h_counts[: 2 * self.world].copy_(self._d_allc, non_blocking=True)
            self._ev_counts.record(main)
            if tr is not None:
                tr.append(("enqueued", time.perf_counter()))
            # the records this turn is expected to add start crossing before their count is known
            ahead_rows = 0
            if host is not None and not overflow:
                ahead_rows = min(hint + 4 * int(hint ** 0.5) + 64, n_samples - n_local_written,
                                 host_cap - n_local_written)
                if ahead_rows > 0:
                    cs.wait_event(self._ev_counts)
                    copy_rows(n_local_written, n_local_written + ahead_rows)
                else:
                    ahead_rows = 0
            if (hint is not None and n_proposed <= max_samples
                    and n_accepted + 2 * hint + 64 < n_samples):
                self.draw_turn(drawsize)  # speculative: overlaps the round trip below
                drawn_ahead = True
            if tr is not None:
                tr.append(("ahead", time.perf_counter()))
            self._ev_counts.synchronize()
            if self.world > 1 and self._peer_exchange() is not None:
                self._peer_exchange().check()
            if tr is not None:
                tr.append(("counts", time.perf_counter()))
            if self.world == 1:
                c_acc, c_written = int(self._h_counts_np[0]), int(self._h_counts_np[1])
                glob = c_acc
            else:
                allc = self._h_counts_np[: 2 * self.world].reshape(self.world, 2)
                c_acc, c_written = int(allc[self.rank, 0]), int(allc[self.rank, 1])
                glob = int(allc[:, 0].sum())
                if shared is not None:
                    # turn-major, rank-major order; the pool keeps the first n_samples records
                    lo = placed + int(allc[: self.rank, 1].sum())
                    hi = min(lo + c_written, n_samples)
                    if hi > lo:
                        self._segments.append((lo, hi, n_local_written))
                        cs.wait_event(self._ev_counts)
                        with torch.cuda.stream(cs):
                            pool_host[lo * rb : hi * rb].copy_(
                                self.d_rows[n_local_written * rb : (n_local_written + hi - lo) * rb],
                                non_blocking=True)
                    placed = min(placed + int(allc[:, 1].sum()), n_samples)
            n_accepted += glob
            hint = glob
            if single and c_written:
                if host is None:
                    host_cap = max(min(n_samples, int(1.3 * c_written * (max_turns - turn + 1)) + 4096), c_written)
                    host = torch.empty(host_cap * rb, dtype=torch.uint8, pin_memory=True)
                if not overflow and n_local_written + c_written <= host_cap:
                    if c_written > ahead_rows:
                        cs.wait_event(self._ev_counts)
                        copy_rows(n_local_written + ahead_rows, n_local_written + c_written)
                    copied = n_local_written + c_written
                else:
                    overflow = True
            n_local_written += c_written
            if tr is not None:
                tr.append(("copy_enqueued", time.perf_counter()))
            if n_proposed > max_samples:
                logger.warning("Reached max samples (%s)", max_samples)
                break
            if n_accepted >= n_samples:
                break
        self._accept_hint = hint
        if drawn_ahead:
            # a draw that was not needed: nothing of it is read, and the likelihood calls its
            # _after_draw made are not counted (the reference counts the turns it uses)
            if self.likelihood is not None and self._last is not None:
                self.likelihood_evaluations = getattr(self, "likelihood_evaluations", 0) - self._last[0]
            self._last = None
        if host is not None or shared is not None:
            self._copy_stream.synchronize()  # d_rows is free for the next populate
        if tr is not None:
            tr.append(("copies_done", time.perf_counter()))
        if not to_host:
            rows = n_local_written
        elif shared is not None:
            shared.barrier()  # every rank's records have landed
            rows = pool_host[: placed * rb].numpy().view(self.row_dtype)
        elif host is not None and not overflow and copied == n_local_written:
            rows = host[: n_local_written * self.row_bytes].numpy().view(self.row_dtype)
        else:
            rows = self._gather_rows(n_local_written, n_samples)
        if tr is not None:
            tr.append(("rows", time.perf_counter()))
        return rows, n_proposed, n_accepted

    # ------------------------------------------------------- accumulate_weights
    def run_accumulate(self, n_samples: int, drawsize: int, max_samples: int = 1_000_000):
        """The ``accumulate_weights=True`` variant of the loop
        (/root/reference/src/nessai/proposal/flowproposal/flowproposal.py:414-417,471-490,504-512):
        the rows and weights of EVERY turn are kept (turn ``t`` is drawn into slot ``t`` of the
        device buffers by the same fused kernel, the running maximum ``log_constant`` and the
        valid count accumulate in ``d_stats``), the expected pool size
        ``exp(logsumexp(log_weights - log_constant))`` is one reduction kernel per turn, and the
        rejection step runs over all rows so far -- fresh uniforms each time -- only once that
        reaches ``n_samples``; ``samples[accept][:n_samples]`` is what the accept kernel's
        capacity implements.  One host synchronisation per turn (this variant is not pipelined).
        Returns ``(rows, n_proposed, n_accepted)`` like ``run``."""
        import math

        if self.log_prior_const is None:
            raise NotImplementedError("nessai_b200: accumulate_weights needs the prior on the device")
        n_samples, drawsize = int(n_samples), int(drawsize)
        ctl = AccumulateControl(n_samples, max_samples)
        n_local, start = self._shard(drawsize)
        # rows per turn slot: the same on every rank, a multiple of 32 rows (vector stores)
        stride = -(-(-(-drawsize // self.world)) // 32) * 32
        max_turns = int(max_samples) // max(drawsize, 1) + 1
        cap = stride * max_turns
        limit = int(os.environ.get("NB200_ACCUMULATE_MAX_BYTES", 64 << 30))
        if cap * (4 * self.D + 16) > limit:
            raise MemoryError(
                f"nessai_b200: accumulate_weights would keep {cap} rows on the device "
                f"({cap * (4 * self.D + 16) / 2**30:.1f} GiB); lower max_samples"
            )
        self._ensure(cap, max(n_samples, 1), False)
        self.model._ready()
        lib = _lib.load()
        dev = self.device
        st = torch.cuda.current_stream(dev).cuda_stream
        if getattr(self, "d_partials", None) is None:
            self.d_partials = torch.empty(self._N_PARTIALS, dtype=torch.float64, device=dev)
        self.d_logw[:cap].fill_(float("nan"))  # slot padding and undrawn slots never count
        self.d_stats.copy_(self._stats_init, non_blocking=True)
        lpc = float(self.log_prior_const)
        info = dict(stride=stride, n_local=n_local, start=start, draw_offsets=[], rejects=[], n_expected=[])
        self.last_accumulate = info
        turn, n_valid = 0, 0
        while ctl.go_on():
            off = turn * stride
            info["draw_offsets"].append(self._turn_rows + start)
            self._call(lib.nb200_populate_draw, [
                self.model._handle, n_local, self._seed(), self._turn_rows + start, self.r_max, self.sqrt_t,
                self.d_scale.data_ptr(), self.d_shift.data_ptr(), self.d_lo.data_ptr(), self.d_hi.data_ptr(),
                lpc, self.min_log_q, self.d_xp.data_ptr() + off * self.D * 4, self.d_logq.data_ptr() + off * 8,
                self.d_logw.data_ptr() + off * 8, None, self._accumulate_draw_stats().data_ptr(), st,
            ], "nb200_populate_draw")
            self._accumulate_after_draw(off, n_local, st)
            self._turn_rows += drawsize
            turn += 1
            rows = turn * stride
            if self.likelihood is not None and n_local > 0:
                # LikelihoodThresholdTruncation inside the accumulating loop (flowproposal.py:456-467):
                # the slot's rows at or below the contour leave, the running statistics are redone
                self.likelihood_evaluations = getattr(self, "likelihood_evaluations", 0) + n_local
                if getattr(self, "d_logl", None) is None or self.d_logl.shape[0] < self._cap:
                    self.d_logl = torch.empty(self._cap, dtype=torch.float64, device=dev)
                    self._gen += 1
                self._likelihood_cut(self._slot_x(off, n_local), off, n_local, rows)
            if self.world > 1:
                import torch.distributed as dist

                dist.all_reduce(self.d_stats[0:1], op=dist.ReduceOp.MAX, group=self.group)
            self._call(lib.nb200_sum_exp, [
                self.d_logw.data_ptr(), rows, self.d_stats.data_ptr(), self.d_partials.data_ptr(),
                self._N_PARTIALS, st,
            ], "nb200_sum_exp")
            # {sum exp(log_w - log_constant), valid rows}: the partials are added in index order
            both = torch.stack([self.d_partials.sum(), self.d_stats[1]])
            if self.world > 1:
                dist.all_reduce(both, op=dist.ReduceOp.SUM, group=self.group)
            n_expected, nv = (float(v) for v in both.cpu())  # the synchronisation of this turn
            info["n_expected"].append(n_expected)
            added = int(nv) > n_valid
            n_valid = int(nv)
            if ctl.turn_drawn(drawsize, added, n_expected):
                ctl.rejected(self._accumulate_reject(rows, n_samples, info))
            ctl.end_turn()
        if ctl.stopped_on_max_samples:
            logger.warning("Reached max samples (%s)", max_samples)
        if turn and ctl.stale:
            ctl.rejected(self._accumulate_reject(turn * stride, n_samples, info))
        n_written = int(self.d_counts.cpu()[1]) if info["rejects"] else 0
        return self._gather_rows(n_written, n_samples), ctl.n_proposed, ctl.n_accepted

    _N_PARTIALS = 296  # blocks of the sum-exp reduction: two per SM

    def _accumulate_draw_stats(self) -> torch.Tensor:
        """Where the draw kernel accumulates {max log_w, n_valid} in the accumulating loop."""
        return self.d_stats

    def _accumulate_after_draw(self, off: int, n_local: int, stream) -> None:
        """Hook after the draw of a turn into slot offset ``off`` (GeneralPopulateEngine: the tail)."""

    def _accumulate_reject(self, rows: int, n_samples: int, info: dict) -> int:
        """Rejection step over the first ``rows`` slot rows with the running maximum
        (flowproposal.py:483-485,505-508); returns the global number of accepted rows.  Every call
        reserves its own block of Philox counters, so the uniforms are fresh."""
        base = self._turn_rows + self.rank * rows
        self._turn_rows += self.world * rows
        info["rejects"].append((base, rows))
        self._accumulate_accept(rows, base, int(n_samples))
        tot = self.d_counts[0:1].clone()
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
        return int(tot.cpu())

    def _slot_x(self, off: int, n: int) -> torch.Tensor:
        """Physical x (float64) of ``n`` rows at row offset ``off`` of the accumulating buffers."""
        return self.d_xp[off : off + n].to(torch.float64) * self.d_scale + self.d_shift

    def _accumulate_logl_ptr(self):
        with_logl = self.likelihood is not None and getattr(self, "d_logl", None) is not None and self.logl_offset >= 0
        return self.d_logl.data_ptr() if with_logl else None

    def _accumulate_accept(self, rows: int, base: int, n_samples: int) -> None:
        self._call(_lib.load().nb200_populate_accept, [
            rows, self.D, self.d_xp.data_ptr(), self.d_scale.data_ptr(), self.d_shift.data_ptr(),
            self.d_logw.data_ptr(), self._accumulate_logl_ptr(), self.d_stats.data_ptr(), self._seed(), base,
            float(self.log_prior_const), self.d_template.data_ptr(), self.row_bytes,
            self.field_offsets.ctypes.data, self.logl_offset, self.d_rows.data_ptr(), n_samples, 0,
            self.d_counts.data_ptr(), self.d_scratch.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream,
        ], "nb200_populate_accept")

    def _shared_pool(self, nbytes: int):
        """The node-local shared host pool (hostpool.py), created collectively on first use and
        whenever a larger one is needed; None when the ranks span several nodes, when it is
        disabled (NB200_NO_SHM=1) or when any rank fails to map it."""
        from .hostpool import SharedHostPool, same_node

        if getattr(self, "_pool_ok", None) is None:
            self._pool_ok = os.environ.get("NB200_NO_SHM", "0") != "1" and same_node(self.group)
            self._pool, self._pool_turn = None, 0
        if not self._pool_ok:
            return None
        if self._pool is None or self._pool.nbytes < nbytes:
            try:  # the constructor fails on every rank or on none
                if self._pool is not None:
                    self._pool.close()
                self._pool = SharedHostPool(nbytes, group=self.group)
            except RuntimeError as e:  # pragma: no cover - depends on the host
                logger.warning("nessai_b200: shared host pool unavailable (%s); using all-gather", e)
                self._pool_ok, self._pool = False, None
                return None
        return self._pool

    def _run_serial(self, n_samples, drawsize, max_samples, host_prior, timing):
        """The same loop, one synchronisation per turn and no overlap (host-side prior, or
        NB200_TIMING phase timing)."""
        self._ensure(1, int(n_samples), False)
        n_proposed = 0
        n_accepted = 0  # global
        n_local_written = 0
        if timing:
            import time

            self.last_phase_s = {"draw": 0.0, "accept": 0.0, "counts": 0.0}

            def tick(name, t0):
                torch.cuda.synchronize(self.device)
                self.last_phase_s[name] += time.perf_counter() - t0
                return time.perf_counter()

        while n_accepted < n_samples:
            if timing:
                torch.cuda.synchronize(self.device)
                tt = time.perf_counter()
            self.draw_turn(int(drawsize))
            if timing:
                tt = tick("draw", tt)
            n_proposed += int(drawsize)
            if host_prior is not None:
                self._apply_host_prior(host_prior)
            counts = self.accept_turn(int(n_samples) - n_local_written, n_local_written)
            if timing:
                tt = tick("accept", tt)
            if self.world > 1:
                import torch.distributed as dist

                tot = counts[0:1].clone()
                dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
                both = torch.cat([counts, tot]).cpu()  # one synchronisation per turn
                c = both[:2]
                n_accepted += int(both[2])
            else:
                c = counts.cpu()
                n_accepted += int(c[0])
            if timing:
                tt = tick("counts", tt)
            n_local_written += int(c[1])
            self._turn_rows += int(drawsize)
            if n_proposed > max_samples:
                logger.warning("Reached max samples (%s)", max_samples)
                break
        if timing:
            torch.cuda.synchronize(self.device)
            t0 = time.perf_counter()
            rows = self._gather_rows(n_local_written, int(n_samples))
            self.last_gather_s = time.perf_counter() - t0
        else:
            rows = self._gather_rows(n_local_written, int(n_samples))
        return rows, n_proposed, n_accepted

    def _apply_host_prior(self, host_prior):
        n_local, _ = self._last
        x = self.physical_x(n_local).cpu().numpy()
        lw = self.d_logw[:n_local].cpu().numpy()
        ok = ~np.isnan(lw)
        xs = empty_structured_array(int(ok.sum()), dtype=self.row_dtype)
        for i, nm in enumerate(self.names):
            xs[nm] = x[ok, i]
        lp = np.asarray(host_prior(xs), dtype=np.float64)
        lw[ok] += lp
        lw[ok & ~np.isfinite(lw)] = np.nan
        self.d_logw[:n_local].copy_(torch.from_numpy(lw))
        good = lw[ok][np.isfinite(lw[ok])]
        self.d_stats[0] = float(good.max()) if good.size else -float("inf")
        self._host_logp = None

    def _gather_rows(self, n_local_written: int, n_samples: int) -> np.ndarray:
        rb = self.row_bytes
        if self.world == 1:
            if not n_local_written:
                return empty_structured_array(0, dtype=self.row_dtype)
            # pinned staging (torch's caching host allocator) -> one async D2H copy; the
            # numpy array aliases the pinned block and keeps it alive
            nbytes = n_local_written * rb
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            host.copy_(self.d_rows[:nbytes], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return host.numpy().view(self.row_dtype)
        full, _ = gather_records(self.d_rows, n_local_written, n_samples, rb, self.group, to_host=True)
        if not full.numel():
            return empty_structured_array(0, dtype=self.row_dtype)
        return full.numpy().view(self.row_dtype)


class GeneralPopulateEngine(PopulateEngine):
    """``PopulateEngine`` for per-parameter maps ``x = h(a x' + b) * scale + shift`` that are not
    all affine: ``h`` = sigmoid (``RescaleToBounds`` with ``post_rescaling="logit"``), exp
    (``"log"``), log, the normal CDF / quantile function (``"inv_gaussian_cdf"`` /
    ``"gaussian_cdf"``) or ``|.|`` (boundary inversion); as a PRE-rescaling the same functions
    come after the affine map ``a, b`` (``"z-score-logit"``, ``"log-z-score"``, ...) --
    /root/reference/src/nessai/reparameterisations/rescale.py:263-291,570-590,635-660,
    utils/rescaling.py:290-417.

    The fused draw kernel is left exactly as it is: it is given the identity map and no bounds,
    so it leaves the flow output ``x'`` and the flow's own ``log q``; ``nb200_reparam_tail``
    (csrc/reparam_tail.cuh) then forms ``x`` and ``log|J|`` in float64, checks the prior bounds
    and ``min_log_q``, rewrites ``log q`` / ``log w`` and the turn's statistics, and the
    rejection step copies the float64 rows into the records (``nb200_populate_accept_x64``).
    The loop, its pipelining and the multi-GPU exchange are the base class's."""

    MAX_D = 64  # TAIL_MAXD
    N_KINDS = 18  # TAIL_N_KINDS

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.D > self.MAX_D:
            raise NotImplementedError(f"nessai_b200: the non-affine tail supports at most {self.MAX_D} parameters")
        self.d_x64 = None
        self._tail_min_log_q = -float("inf")

    def configure(self, kind, scale, shift, lo, hi, log_prior_const, r_max, sqrt_temperature=1.0, min_log_q=None,
                  likelihood=None, log_l_threshold=None, pre_scale=None, pre_shift=None, src=None):
        """As ``PopulateEngine.configure`` with the per-parameter ``kind`` of ``h`` in front
        (0 identity, 1 sigmoid, 2 abs, 3 exp, 4 log, 5 normal CDF, 6 normal quantile) and the
        optional affine map applied BEFORE ``h``: ``x = h(pre_scale x' + pre_shift) scale + shift``.
        ``src`` (``(D, 3)`` ints; ``(D, 2)`` is padded): the flow feature(s) output slot ``d`` reads
        (default ``d``); the pair kinds of ``Angle`` read two (7 angle, 8 angle mod 2 pi, 9 radius,
        10 auxiliary radius with its chi(2) prior; 12 ``ToCartesian``), the kinds of ``AnglePair``
        three (13 zenith, 14 declination, 15 radius, 16 auxiliary radius with its chi(3) prior);
        11 is ``floor`` (``Dequantise``), 17 an augment parameter of ``AugmentedFlowProposal`` (identity,
        N(0, 1) prior).  See include/nessai_b200.h: nb200_reparam_tail."""
        D = self.D
        src = np.stack([np.arange(D)] * 3, axis=1) if src is None else np.asarray(src)
        if src.ndim != 2:
            src = src.reshape(D, -1)
        if src.shape[1] == 2:
            src = np.concatenate([src, src[:, :1]], axis=1)
        pre_scale = np.ones(D) if pre_scale is None else pre_scale
        pre_shift = np.zeros(D) if pre_shift is None else pre_shift
        if getattr(self, "_identity", None) is None:
            self._identity = (np.ones(D), np.zeros(D), np.full(D, -np.inf), np.full(D, np.inf))
        super().configure(*self._identity, log_prior_const, r_max, sqrt_temperature, min_log_q=None,
                          likelihood=likelihood, log_l_threshold=log_l_threshold)
        self._tail_min_log_q = -float("inf") if min_log_q is None or np.isnan(min_log_q) else float(min_log_q)
        new = [np.array(kind, dtype=np.int32)] + [
            np.array(a, dtype=np.float64) for a in (scale, shift, lo, hi, pre_scale, pre_shift)]
        if any(a.shape != (D,) for a in new):
            raise ValueError("kind / scale / shift / lo / hi / pre_scale / pre_shift: one entry per parameter")
        if np.any((new[0] < 0) | ((new[0] & 0xFF) >= self.N_KINDS) | ((new[0] & ~0x1FF) != 0)):
            raise ValueError("unknown per-parameter map kind")
        new.append(np.ascontiguousarray(src, dtype=np.int32))
        if new[-1].shape != (D, 3) or np.any((new[-1] < 0) | (new[-1] >= D)):
            raise ValueError("src must hold two or three flow-feature indices per parameter")
        old = getattr(self, "_tail_host", None)
        if old is None or not all(np.array_equal(a, b) for a, b in zip(new, old)):
            self.t_kind = torch.from_numpy(new[0]).to(self.device)
            dev = torch.from_numpy(np.stack(new[1:7])).to(self.device)
            self.t_scale, self.t_shift, self.t_lo, self.t_hi, self.t_pre_scale, self.t_pre_shift = dev.unbind(0)
            self.t_src = torch.from_numpy(new[7].reshape(-1)).to(self.device)
            self._tail_host = new

    def _after_draw(self, n_local: int):
        if self.d_x64 is None or self.d_x64.shape[0] < self._cap:
            self.d_x64 = torch.empty((self._cap, self.D), dtype=torch.float64, device=self.device)
        if n_local > 0:
            self.d_stats.copy_(self._stats_init, non_blocking=True)  # the draw kernel's were pre-tail
            lpc = float("nan") if self.log_prior_const is None else float(self.log_prior_const)
            self._call(_lib.load().nb200_reparam_tail, [
                n_local, self.D, self.d_xp.data_ptr(), self.t_kind.data_ptr(), self.t_src.data_ptr(),
                self.t_pre_scale.data_ptr(),
                self.t_pre_shift.data_ptr(), self.t_scale.data_ptr(), self.t_shift.data_ptr(),
                self.t_lo.data_ptr(), self.t_hi.data_ptr(), lpc, self._tail_min_log_q,
                self.d_logq.data_ptr(), self.d_logw.data_ptr(), self.d_x64.data_ptr(), self.d_stats.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream,
            ], "nb200_reparam_tail")
        super()._after_draw(n_local)

    def physical_x(self, n: int) -> torch.Tensor:
        return self.d_x64[:n]

    def accept_turn(self, capacity_left: int, write_offset: int):
        n_local, start = self._last
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(self.d_stats[0:1], op=dist.ReduceOp.MAX, group=self.group)
        with_logl = self.likelihood is not None and getattr(self, "d_logl", None) is not None and self.logl_offset >= 0
        lp = 0.0 if self.log_prior_const is None else float(self.log_prior_const)
        self._call(_lib.load().nb200_populate_accept_x64, [
            n_local, self.D, self.d_x64.data_ptr(), self.d_logw.data_ptr(),
            self.d_logl.data_ptr() if with_logl else None, self.d_stats.data_ptr(), self._seed(),
            self._turn_rows + start, lp, self.d_template.data_ptr(), self.row_bytes,
            self.field_offsets.ctypes.data, self.logl_offset, self.d_rows.data_ptr(), int(capacity_left),
            int(write_offset), self.d_counts.data_ptr(), self.d_scratch.data_ptr(),
            torch.cuda.current_stream(self.device).cuda_stream,
        ], "nb200_populate_accept_x64")
        return self.d_counts

    # accumulate_weights: the base loop with the tail applied to each turn's slot; the draw kernel's
    # own (pre-tail) statistics go to a scratch pair, the tail's accumulate in d_stats
    def _accumulate_draw_stats(self) -> torch.Tensor:
        if getattr(self, "d_stats_scratch", None) is None:
            self.d_stats_scratch = torch.empty(2, dtype=torch.float64, device=self.device)
        self.d_stats_scratch.copy_(self._stats_init, non_blocking=True)
        return self.d_stats_scratch

    def _accumulate_after_draw(self, off: int, n_local: int, stream) -> None:
        if self.d_x64 is None or self.d_x64.shape[0] < self._cap:
            self.d_x64 = torch.empty((self._cap, self.D), dtype=torch.float64, device=self.device)
        if n_local <= 0:
            return
        self._call(_lib.load().nb200_reparam_tail, [
            n_local, self.D, self.d_xp.data_ptr() + off * self.D * 4, self.t_kind.data_ptr(), self.t_src.data_ptr(),
            self.t_pre_scale.data_ptr(), self.t_pre_shift.data_ptr(), self.t_scale.data_ptr(), self.t_shift.data_ptr(),
            self.t_lo.data_ptr(), self.t_hi.data_ptr(), float(self.log_prior_const), self._tail_min_log_q,
            self.d_logq.data_ptr() + off * 8, self.d_logw.data_ptr() + off * 8,
            self.d_x64.data_ptr() + off * self.D * 8, self.d_stats.data_ptr(), stream,
        ], "nb200_reparam_tail")

    def _slot_x(self, off: int, n: int) -> torch.Tensor:
        return self.d_x64[off : off + n]

    def _accumulate_accept(self, rows: int, base: int, n_samples: int) -> None:
        self._call(_lib.load().nb200_populate_accept_x64, [
            rows, self.D, self.d_x64.data_ptr(), self.d_logw.data_ptr(), self._accumulate_logl_ptr(),
            self.d_stats.data_ptr(), self._seed(), base, float(self.log_prior_const), self.d_template.data_ptr(), self.row_bytes,
            self.field_offsets.ctypes.data, self.logl_offset, self.d_rows.data_ptr(), n_samples, 0,
            self.d_counts.data_ptr(), self.d_scratch.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream,
        ], "nb200_populate_accept_x64")


class B200FlowProposal:
    """Standalone mirror of ``FlowProposal`` for the hot path (see module doc).

    ``model`` needs ``names``, ``bounds`` (name -> (lo, hi)), ``log_prior(x)``
    and ``log_likelihood(x)`` on structured arrays -- the surface of
    /root/reference/src/nessai/model.py:53 that this path touches.
    """

    _FlowModelClass = B200FlowModel

    def __init__(
        self,
        model,
        rng: Optional[np.random.Generator] = None,
        flow_config=None,
        training_config=None,
        output=None,
        poolsize=None,
        drawsize=None,
        plot=False,
        check_acceptance=False,
        max_poolsize_scale=10,
        update_poolsize=True,
        accumulate_weights=False,
        fallback_reparameterisation="zscore",
        latent_temperature=None,
        volume_fraction=0.95,
        fixed_radius=None,
        device_prior="auto",
        truncation_methods=None,
    ):
        methods = ["latent_radius"] if truncation_methods is None else list(dict.fromkeys(truncation_methods))
        unknown = set(methods) - {"latent_radius", "min_log_q", "likelihood_threshold"}
        if unknown:
            # truncation.py:431-435 (TRUNCATION_REGISTRY)
            raise ValueError(f"Unknown truncation method(s): {sorted(unknown)}")
        self.truncation_methods = methods
        if fallback_reparameterisation not in ("zscore", "null", None):
            raise NotImplementedError(
                "nessai_b200: only the 'zscore' and 'null' reparameterisations run on the device"
            )
        self.model = model
        self.rng = rng if rng is not None else np.random.default_rng()
        self.flow_config = dict(flow_config or {})
        self.training_config = dict(training_config or {})
        self.output = output if output is not None else os.getcwd()
        self._poolsize = 10000 if poolsize is None else int(poolsize)
        self._poolsize_scale = 1.0
        self.max_poolsize_scale = max_poolsize_scale
        self.update_poolsize = update_poolsize
        self.drawsize = self._poolsize if drawsize is None else int(drawsize)
        self.check_acceptance = check_acceptance
        self.accumulate_weights = bool(accumulate_weights)
        self.fallback_reparameterisation = fallback_reparameterisation
        if latent_temperature is not None:
            if isinstance(latent_temperature, bool) or not isinstance(latent_temperature, (int, float)):
                raise TypeError("latent_temperature must be a float")
            if latent_temperature <= 0.0:
                raise ValueError("latent_temperature must be positive")
            latent_temperature = float(latent_temperature)
        self.latent_temperature = latent_temperature
        self.volume_fraction = volume_fraction
        self.fixed_radius = fixed_radius
        self.device_prior = device_prior
        self.flow = None
        self.initialised = False
        self.populated = False
        self.populating = False
        self.indices = []
        self.samples = None
        self.x = None
        self.training_count = 0
        self.populated_count = 0
        self.population_acceptance = None
        self.population_time = datetime.timedelta()
        self.ns_acceptance = 1.0
        self.acceptance = []
        self.r = np.nan
        self._checked_population = True
        self.training_data = None
        self._engine = None
        self._min_log_q, self._loop_likelihood, self._log_l_threshold = None, None, None
        self.names = list(model.names)
        self.prime_parameters = [f"{n}_prime" if fallback_reparameterisation == "zscore" else n for n in self.names]
        self.scale = np.ones(len(self.names))
        self.shift = np.zeros(len(self.names))

    # ------------------------------------------------------------- properties
    @property
    def poolsize(self):
        return int(self._poolsize_scale * self._poolsize)

    @property
    def dims(self):
        return len(self.names)

    prime_dims = dims

    @property
    def x_dtype(self):
        return get_dtype(self.names)

    population_dtype = x_dtype

    def update_poolsize_scale(self, acceptance):
        """flowproposal/base.py:416-435."""
        if not acceptance:
            self._poolsize_scale = self.max_poolsize_scale
        else:
            self._poolsize_scale = min(max(1.0 / acceptance, 1.0), self.max_poolsize_scale)

    # ------------------------------------------------------------- lifecycle
    def initialise(self, resumed: bool = False) -> None:
        """flowproposal/base.py:358-391."""
        os.makedirs(self.output, exist_ok=True)
        self.flow_config["n_inputs"] = self.dims
        self.flow = self._FlowModelClass(
            flow_config=self.flow_config,
            training_config=self.training_config,
            output=self.output,
            rng=self.rng,
        )
        self.flow.initialise()
        if self.fixed_radius:
            self.radius = float(self.fixed_radius)
        else:
            self.radius = float(compute_radius(self.dims, self.volume_fraction))
        self.populated = False
        self.initialised = True

    # -------------------------------------------------------------- rescaling
    def check_state(self, x):
        """z-score statistics from the training set
        (reparameterisations/rescale.py:293-304: np.std / np.mean)."""
        if self.fallback_reparameterisation == "zscore":
            self.scale = np.array([np.std(x[n]) for n in self.names])
            self.shift = np.array([np.mean(x[n]) for n in self.names])

    def rescale(self, x, **kwargs):
        """x -> (x', log|J|)  (flowproposal/base.py:716-753, rescale.py:233-261)."""
        x = np.atleast_1d(x)
        x_prime = empty_structured_array(x.size, self.prime_parameters)
        log_J = np.zeros(x.size)
        for i, (n, pn) in enumerate(zip(self.names, self.prime_parameters)):
            x_prime[pn] = (x[n] - self.shift[i]) / self.scale[i]
            log_J -= np.log(np.abs(self.scale[i]))
        for p in NON_SAMPLING_PARAMETERS:
            x_prime[p] = x[p]
        return x_prime, log_J

    def inverse_rescale(self, x_prime, **kwargs):
        """x' -> (x, log|J|)  (flowproposal/base.py:755-784, rescale.py:263-291)."""
        x = empty_structured_array(x_prime.size, self.names)
        log_J = np.zeros(x.size)
        for i, (n, pn) in enumerate(zip(self.names, self.prime_parameters)):
            x[n] = x_prime[pn] * self.scale[i] + self.shift[i]
            log_J += np.log(np.abs(self.scale[i]))
        for p in NON_SAMPLING_PARAMETERS:
            x[p] = x_prime[p]
        return x, log_J

    # ---------------------------------------------------------------- training
    def train(self, x, plot=False):
        """flowproposal/base.py:870-925."""
        if not self.initialised:
            raise RuntimeError("B200FlowProposal is not initialised.")
        self.training_data = x.copy()
        self.check_state(self.training_data)
        x_prime, _ = self.rescale(x)
        self.training_data_prime = x_prime.copy()
        x_prime_array = live_points_to_array(x_prime, self.prime_parameters, copy=True)
        self.flow.train(x_prime_array, output=self.output, plot=False)
        self.populated = False
        self.training_count += 1

    def reset_model_weights(self, **kwargs):
        self.flow.reset_model(**kwargs)

    # ----------------------------------------------------------------- passes
    def forward_pass(self, x, rescale=True, **kwargs):
        """flowproposal/base.py:961-994."""
        log_J = 0
        if rescale:
            x, log_J_rescale = self.rescale(x, **kwargs)
            log_J += log_J_rescale
        names = self.prime_parameters if rescale else [n for n in x.dtype.names if n not in NON_SAMPLING_PARAMETERS]
        x = live_points_to_array(x, names=names, copy=True)
        if x.ndim == 1:
            x = x[np.newaxis, :]
        z, log_prob = self.flow.forward_and_log_prob(x)
        return z, log_prob + log_J

    def latent_log_prob(self, z, temperature=None):
        """flowproposal/base.py:401-414."""
        z = np.asarray(z)
        if temperature in (None, 1.0):
            z_in, log_j = z, 0.0
        else:
            scale = np.sqrt(float(temperature))
            z_in, log_j = z / scale, z.shape[-1] * np.log(scale)
        z_tensor = self.flow.numpy_array_to_tensor(z_in)
        log_p = self.flow.model.base_distribution_log_prob(z_tensor)
        return log_p.cpu().numpy().astype(np.float64) - log_j

    def check_prior_bounds(self, x, *args):
        flags = np.ones(x.size, dtype=bool)
        for n in self.names:
            lo, hi = self.model.bounds[n]
            flags &= ~((x[n] < lo) | (x[n] > hi))
        return (a[flags] for a in (x,) + args)

    def backward_pass(self, z, rescale=True, discard_nans=True, return_z=False, **kwargs):
        """flowproposal/flowproposal.py:345-389 (host-orchestrated variant; the
        fused path is ``populate``)."""
        x, log_j = self.flow.inverse(z)
        log_prob = self.latent_log_prob(z, self.latent_temperature) - log_j
        if discard_nans:
            valid = np.isfinite(log_prob)
            x, log_prob, z = x[valid], log_prob[valid], z[valid]
        xs = empty_structured_array(x.shape[0], self.prime_parameters)
        for i, p in enumerate(self.prime_parameters):
            xs[p] = x[:, i]
        if rescale:
            xs, log_J = self.inverse_rescale(xs)
            log_prob -= log_J
            xs, z, log_prob = self.check_prior_bounds(xs, z, log_prob)
        if return_z:
            return xs, log_prob, z
        return xs, log_prob

    def log_prior(self, x):
        return np.asarray(self.model.log_prior(x), dtype=np.float64)

    def compute_weights(self, x, log_q, return_log_prior=False):
        log_p = self.log_prior(x)
        log_w = log_p - log_q
        return (log_w, log_p) if return_log_prior else log_w

    # --------------------------------------------------------------- populate
    def _get_engine(self):
        if self._engine is None or self._engine.flow is not self.flow:
            self._engine = PopulateEngine(self.flow, self.names, self.x_dtype)
            if self.device_prior in ("auto", True, "uniform"):
                self._log_prior_const = detect_uniform_box_prior(self.model, self.rng)
                if self.device_prior in (True, "uniform") and self._log_prior_const is None:
                    raise RuntimeError("device_prior requested but the prior is not a finite uniform box")
            else:
                self._log_prior_const = None
        lo = [self.model.bounds[n][0] for n in self.names]
        hi = [self.model.bounds[n][1] for n in self.names]
        if getattr(self, "_bounds_cache", None) is None or self._bounds_cache[0] != (lo, hi):
            self._bounds_cache = ((lo, hi), np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64))
        lo, hi = self._bounds_cache[1], self._bounds_cache[2]
        t = self.latent_temperature
        self._engine.configure(
            self.scale, self.shift, lo, hi, self._log_prior_const,
            self.radius if "latent_radius" in self.truncation_methods else 0.0,
            1.0 if t in (None, 1.0) else float(np.sqrt(t)),
            min_log_q=self._min_log_q, likelihood=self._loop_likelihood, log_l_threshold=self._log_l_threshold,
        )
        return self._engine

    def _prepare_truncation(self, worst_point):
        """``TruncationScheme.prepare`` for the optional rules (truncation.py:378-386,410-420)."""
        self._min_log_q, self._loop_likelihood, self._log_l_threshold = None, None, None
        if "min_log_q" in self.truncation_methods:
            if self.training_data is None or not len(self.training_data):
                raise RuntimeError("min_log_q truncation requires training_data to be set")
            self._min_log_q = float(self.forward_pass(self.training_data)[1].min())
        if "likelihood_threshold" in self.truncation_methods:
            fn = getattr(self.model, "log_likelihood_torch", None)
            if fn is None:
                raise NotImplementedError(
                    "nessai_b200: likelihood_threshold truncation runs inside the device loop and needs "
                    "model.log_likelihood_torch(x: (n, D) float64 device tensor) -> (n,) tensor"
                )
            thr = np.asarray(worst_point["logL"], dtype=float).reshape(-1)[0] if worst_point is not None else np.nan
            self._loop_likelihood = fn
            self._log_l_threshold = float(thr) if np.isfinite(thr) else -np.inf

    def populate(self, worst_point, n_samples=10000, plot=False, r=None, max_samples=1_000_000) -> None:
        """flowproposal/flowproposal.py:391-534 with the loop body on the GPU."""
        st = datetime.datetime.now()
        if not self.initialised:
            raise RuntimeError(
                "Proposal has not been initialised. Try calling `initialise()` first."
            )
        if r is not None:
            self.radius = float(r)
        self.indices = []
        self._prepare_truncation(worst_point)
        eng = self._get_engine()
        host_prior = None if self._log_prior_const is not None else self.log_prior
        if self.accumulate_weights:
            # flowproposal.py:471-490,504-512 (raises if the prior is not on the device)
            rows, n_proposed, n_accepted = eng.run_accumulate(
                int(n_samples), int(self.drawsize), max_samples=max_samples
            )
        else:
            rows, n_proposed, n_accepted = eng.run(
                int(n_samples), int(self.drawsize), max_samples=max_samples, host_prior=host_prior
            )
        self.x = rows
        self.samples = rows
        if host_prior is not None and len(rows):
            self.samples["logP"] = self.log_prior(self.samples)
        self.n_proposed = n_proposed
        self.population_time += datetime.datetime.now() - st
        if len(self.samples) and self._loop_likelihood is None:
            # flowproposal.py:519-523; on the device when the model offers it (the accepted
            # records are still resident, so only 8 bytes per row come back)
            fn = getattr(self.model, "log_likelihood_torch", None)
            if fn is not None and eng.world == 1:
                self.samples["logL"] = eng.device_log_likelihood(len(rows), fn).cpu().numpy()
            elif fn is not None and host_prior is None and eng.sharded_pool_likelihood(self.samples, fn):
                pass  # several GPUs: every rank evaluated the records it accepted itself
            else:
                self.samples["logL"] = np.asarray(self.model.log_likelihood(self.samples))
        if self.check_acceptance:
            self.acceptance.append(self.compute_acceptance(worst_point["logL"]))
        self.indices = IndexPool(self.rng.permutation(self.samples.size))
        self.population_acceptance = n_accepted / n_proposed
        self.populated_count += 1
        self.populated = True
        self._checked_population = False

    def compute_acceptance(self, logL):
        return (self.samples["logL"] > logL).sum() / self.samples.size

    def draw(self, worst_point):
        """flowproposal/base.py:1152-1183."""
        if not self.populated:
            self.populating = True
            if self.update_poolsize:
                self.update_poolsize_scale(self.ns_acceptance)
            while not self.populated:
                self.populate(worst_point, n_samples=self.poolsize)
            self.populating = False
        index = self.indices.pop()
        new_sample = self.samples[index]
        if not self.indices:
            self.populated = False
        return new_sample

    def reset(self):
        self.indices = []
        self.samples = None
        self.x = None
        self.populated = False
        self.populated_count = 0
        self.population_acceptance = None
        self._poolsize_scale = 1.0
        self._checked_population = True
        self.acceptance = []

    def __getstate__(self):
        state = self.__dict__.copy()
        state["initialised"] = False
        state["weights_file"] = getattr(state.get("flow"), "weights_file", None)
        state["resume_populated"] = bool(state["populated"] and state["indices"])
        for k in ("model", "flow", "_engine", "_loop_likelihood"):
            state.pop(k, None)
        return state

    def resume(self, model, flow_config, weights_file=None):
        """flowproposal/base.py:1237-1271."""
        self.model = model
        self.flow_config = dict(flow_config)
        self._engine = None
        self.initialise(resumed=True)
        if weights_file is None:
            weights_file = getattr(self, "weights_file", None)
        if weights_file is not None and os.path.exists(weights_file):
            self.flow.reload_weights(weights_file)
