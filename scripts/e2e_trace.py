"""Host timeline of B200FlowProposal.populate (NB200_TRACE): where the end-to-end time goes.
One GPU: python scripts/e2e_trace.py;  N GPUs: torchrun --nproc-per-node N scripts/e2e_trace.py
(weak scaling like bench.py: 1e6 rows per GPU per turn).  Medians over 40 populates, rank 0."""
import os, sys, time, tempfile, collections
os.environ["NB200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import bench
from nessai_b200.livepoint import numpy_array_to_live_points

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
live, _ = bench.live_points()
model = bench.GaussianModel()
pool = 1_000_000 * world
live_s = numpy_array_to_live_points(live, model.names)
live_s["logL"] = model.log_likelihood(live_s)
worst = live_s[np.argmin(live_s["logL"])]
prop = bench.build_proposal("c2_realnvp_mlp", model, live_s, local_rank, pool)
for _ in range(5):
    prop.populate(worst, n_samples=pool, max_samples=pool)
seg = collections.OrderedDict(); tot = []; pre = []; post = []
for rep in range(40):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    prop.population_time *= 0
    t0 = time.perf_counter()
    prop.populate(worst, n_samples=pool, max_samples=pool)
    t1 = time.perf_counter()
    tr = prop._engine.last_trace
    tot.append(prop.population_time.total_seconds() * 1e3); pre.append((tr[0][1] - t0) * 1e3)
    k = 0
    for (a, ta), (b, tb) in zip(tr[:-1], tr[1:]):
        seg.setdefault((k, a, b), []).append(1e3 * (tb - ta)); k += 1
if rank == 0:
    print(f"world {world}: population_time median {np.median(tot):.3f} ms (IQR {np.subtract(*np.percentile(tot, [75, 25])):.3f}); "
          f"run() entered at +{np.median(pre):.3f} ms")
    for (k, a, b), v in seg.items():
        print(f"   {a:>14s} -> {b:<14s} {np.median(v):7.3f} ms   (p90 {np.percentile(v, 90):.3f})")
if world > 1:
    dist.destroy_process_group()
