"""Node-local host memory shared by the ranks of a multi-GPU populate.

Every rank of an SPMD run needs the whole pool of accepted live points on its host
(each rank drives an identical sampler; the reference's API returns the pool as a numpy
array, /root/reference/src/nessai/proposal/flowproposal/flowproposal.py:510-516).  With one
private host copy per rank the records cross PCIe ``world`` times -- at 8 GPUs that is what
bounds the end-to-end rate, not the kernels.  Here the ranks of a node map ONE block of
POSIX shared memory (page-locked with ``cudaHostRegister`` so device-to-host copies stay
asynchronous); each rank copies only the records it accepted itself, at the offsets the
exchanged counts assign, and after a host-side barrier every rank reads the same bytes.

The block holds ``n_buffers`` pools used round-robin: an array handed out by populate k
is overwritten by populate ``k + n_buffers`` (the sampler replaces ``proposal.samples`` on
every populate, so two is enough; a caller that keeps older pools must copy them).
"""

from __future__ import annotations

import os
import socket
import time
import uuid

import numpy as np
import torch

_PAGE = 4096


def same_node(group=None) -> bool:
    """True if every rank of ``group`` runs on this host."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    names = [None] * world
    dist.all_gather_object(names, socket.gethostname(), group=group)
    return len(set(names)) == 1


class SharedHostPool:
    def __init__(self, nbytes: int, group=None, n_buffers: int = 2, register: bool = True):
        import torch.distributed as dist

        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = (int(nbytes) + _PAGE - 1) // _PAGE * _PAGE
        self.n_buffers = int(n_buffers)
        total = _PAGE + self.n_buffers * self.nbytes
        src = dist.get_global_rank(group, 0) if group is not None else 0
        box = [None]
        if self.rank == 0:
            try:
                path = f"/dev/shm/nb200-{os.getpid()}-{uuid.uuid4().hex[:12]}"
                with open(path, "wb") as f:
                    f.truncate(total)
                box[0] = path
            except OSError:
                box[0] = None
        dist.broadcast_object_list(box, src=src, group=group)
        if box[0] is None:
            raise RuntimeError("cannot create a shared-memory block under /dev/shm")
        try:
            self._map = np.memmap(box[0], dtype=np.uint8, mode="r+", shape=(total,))
            mapped = True
        except (OSError, ValueError):
            mapped = False
        oks = [None] * self.world
        dist.all_gather_object(oks, mapped, group=group)  # also: every rank has opened the file
        if self.rank == 0:
            os.unlink(box[0])  # the mappings keep the memory alive; nothing is left behind
        if not all(oks):
            raise RuntimeError("a rank could not map the shared-memory block")
        self._ctrl = self._map[:_PAGE].view(np.int64)
        self._epoch = 0
        self.registered = False
        if register and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self._map.ctypes.data, total, 0)
            self.registered = int(rc) == 0
        self.arrays = [
            self._map[_PAGE + k * self.nbytes : _PAGE + (k + 1) * self.nbytes] for k in range(self.n_buffers)
        ]
        self.tensors = [torch.from_numpy(a) for a in self.arrays]

    def barrier(self, timeout: float = 300.0) -> None:
        """Host-side barrier over the ranks of the node (epoch counters in the shared block)."""
        self._epoch += 1
        self._ctrl[self.rank] = self._epoch
        seen = self._ctrl[: self.world]
        t0 = time.monotonic()
        while int(seen.min()) < self._epoch:
            if time.monotonic() - t0 > timeout:
                raise RuntimeError("nessai_b200: timed out waiting for the other ranks at the host-pool barrier")

    def close(self) -> None:
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self._map.ctypes.data)
            self.registered = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
