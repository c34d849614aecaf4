"""Peer-memory exchange between the ranks of ONE node (csrc/xchg.cuh, C ABI ``nb200_xchg_*``).

The ranks of a sharded populate turn exchange a few bytes twice per turn: the turn's max
log-weight before the rejection step (/root/reference/src/nessai/proposal/flowproposal/
flowproposal.py:491-494 normalises by the maximum over the whole turn) and every rank's
``{accepted, written}`` counts after it.  As NCCL collectives those are two launch- and
protocol-latency-bound calls on the critical path of a ~300 us turn; here each rank owns a
small buffer of slots that its peers write with direct NVLink stores (CUDA IPC mapping) and
that a one-warp kernel polls locally -- no collective library, no host round trip.
"""

from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


class PeerExchange:
    KIND_MAX, KIND_COUNTS = 0, 1

    def __init__(self, device: torch.device, group=None):
        import torch.distributed as dist

        self.device = torch.device(device)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lib = _lib.load()
        self._buf = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self._peers = []
        ok = True
        try:
            with torch.cuda.device(self.device):
                _lib.check(lib.nb200_xchg_create(C.byref(self._buf), handle), "nb200_xchg_create")
        except Exception:
            ok = False
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if ok else None, group=group)
        ok = ok and all(h is not None for h in handles)
        ptrs = []
        if ok:
            try:
                with torch.cuda.device(self.device):
                    for r, h in enumerate(handles):
                        if r == self.rank:
                            ptrs.append(self._buf.value)
                            continue
                        p = C.c_void_p()
                        hb = (C.c_ubyte * 64).from_buffer_copy(h)
                        _lib.check(lib.nb200_xchg_open(hb, C.byref(p)), "nb200_xchg_open")
                        self._peers.append(p)
                        ptrs.append(p.value)
            except Exception:
                ok = False
        oks = [None] * self.world
        dist.all_gather_object(oks, ok, group=group)  # also: every rank has mapped every buffer
        self.ok = all(oks)
        if not self.ok:
            self.close()
            return
        self.d_peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.d_err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.h_err = torch.zeros(1, dtype=torch.int32, pin_memory=True)
        self.seq = [0, 0]
        self._fn = lib.nb200_xchg_allgather

    def allgather(self, kind: int, src: torch.Tensor, n_words: int, gathered=None, max_out=None) -> None:
        """Publish ``n_words`` 64-bit words of ``src`` (device) to every rank; ``gathered``
        (device, ``world * n_words`` 64-bit words) receives every rank's, rank-major; ``max_out``
        (device float64) the maximum over the ranks of word 0 read as a double.  Stream-ordered."""
        self.seq[kind] += 1
        rc = self._fn(
            self.d_peers.data_ptr(), self.world, self.rank, kind, self.seq[kind], src.data_ptr(), n_words,
            None if gathered is None else gathered.data_ptr(), None if max_out is None else max_out.data_ptr(),
            self.d_err.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream,
        )
        if rc:
            _lib.check(rc, "nb200_xchg_allgather")

    def stage_error_flag(self) -> None:
        """Queue the copy of the time-out flag next to the caller's own read-back."""
        self.h_err.copy_(self.d_err, non_blocking=True)

    def check(self) -> None:
        """After the caller synchronised: raise if a peer did not answer an exchange."""
        if int(self.h_err[0]):
            raise RuntimeError("nessai_b200: a rank did not answer a peer-memory exchange (time-out)")

    def close(self) -> None:
        lib = _lib.load()
        for p in self._peers:
            lib.nb200_xchg_close(p)
        self._peers = []
        if self._buf.value:
            lib.nb200_xchg_destroy(self._buf)
            self._buf = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_peer_exchange(device: torch.device, group=None):
    """A ``PeerExchange`` for the ranks of ``group`` if they share a node and CUDA IPC works
    between them (collective: every rank calls it), else ``None`` (NCCL collectives are used).
    ``NB200_NO_XCHG=1`` disables it."""
    import torch.distributed as dist

    from .hostpool import same_node

    if torch.device(device).type != "cuda" or os.environ.get("NB200_NO_XCHG", "0") == "1":
        return None
    if dist.get_world_size(group) > 32 or not same_node(group):
        return None
    x = PeerExchange(device, group)
    return x if x.ok else None
