"""Host-side logic of the multi-GPU populate, on CPU with world_size=2 over gloo:
sharding of a turn, the scalar exchanges and the all-gather of accepted records."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nessai_b200.livepoint import get_dtype
    from nessai_b200.proposal import gather_records, shard_rows

    dtype = get_dtype(["a", "b"])
    rb = dtype.itemsize
    n_total = 1001
    n_local, first = shard_rows(n_total, rank, world)
    # every rank "accepts" the rows whose global index is a multiple of 3 (+ rank-specific count)
    idx = np.arange(first, first + n_local)
    keep = idx[idx % 3 == 0]
    rows = np.zeros(len(keep), dtype=dtype)
    rows["a"] = keep
    rows["b"] = -keep.astype(float)
    rows["it"] = rank
    buf = torch.from_numpy(rows.view(np.uint8).copy())
    # the max normaliser and the accepted count are exchanged as scalars
    mx = torch.tensor([float(keep.max())], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    tot = torch.tensor([len(keep)], dtype=torch.int64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    full, counts = gather_records(buf, len(keep), 300, rb)
    got = full.numpy().view(dtype)
    if rank == 0:
        torch.save(dict(a=got["a"].copy(), it=got["it"].copy(), mx=float(mx), tot=int(tot), counts=counts,
                        shards=[shard_rows(n_total, r, world) for r in range(world)]), out)
    dist.destroy_process_group()


def test_shard_and_gather_world2(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    # shards tile the draw exactly
    assert res["shards"] == [(501, 0), (500, 501)]
    expect = np.arange(0, 1001, 3)
    assert res["tot"] == len(expect) and res["mx"] == float(expect.max())
    assert sum(res["counts"]) == len(expect)
    # rank-major == global draw order here (contiguous shards); first 300 kept
    np.testing.assert_array_equal(res["a"], expect[:300])
    assert set(res["it"].tolist()) <= {0, 1}


def _pool_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nessai_b200.hostpool import SharedHostPool, same_node

    assert same_node()
    pool = SharedHostPool(10_000, n_buffers=2, register=False)
    assert pool.nbytes == 12288 and len(pool.tensors) == 2
    seen = []
    for turn in range(3):  # buffers are used round-robin
        buf = pool.tensors[turn % 2]
        lo, hi = (0, 3000) if rank == 0 else (3000, 7000)
        buf[lo:hi] = torch.full((hi - lo,), 10 * turn + rank + 1, dtype=torch.uint8)
        pool.barrier()
        seen.append(buf[:7000].clone())  # every rank reads the bytes of both
        pool.barrier()  # nobody overwrites before everybody has read
    torch.save(seen, f"{out}.{rank}")
    dist.destroy_process_group()


def test_shared_host_pool_world2(tmp_path):
    """hostpool.py: one block of node-local shared memory, each rank writes its own
    records, a host barrier, every rank sees the whole pool."""
    out = str(tmp_path / "pool")
    mp.spawn(_pool_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = (torch.load(f"{out}.{r}", weights_only=False) for r in range(2))
    for turn in range(3):
        assert torch.equal(a[turn], b[turn])
        assert set(a[turn][:3000].tolist()) == {10 * turn + 1} and set(a[turn][3000:].tolist()) == {10 * turn + 2}
    assert not [f for f in os.listdir("/dev/shm") if f.startswith(f"nb200-")]  # unlinked once mapped


def test_shard_rows_properties():
    from nessai_b200.proposal import shard_rows

    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            parts = [shard_rows(n, r, world) for r in range(world)]
            assert sum(p[0] for p in parts) == n
            pos = 0
            for cnt, first in parts:
                assert first == pos
                pos += cnt
            assert max(p[0] for p in parts) - min(p[0] for p in parts) <= 1
