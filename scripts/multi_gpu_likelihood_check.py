"""N ranks: the pool's logL evaluated by every rank on its own accepted records (device, sharded)
equals the host likelihood of the whole pool.  torchrun --nproc-per-node N scripts/multi_gpu_likelihood_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from nessai_b200.livepoint import numpy_array_to_live_points

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model = bench.GaussianModel(); live, _ = bench.live_points()
live_s = numpy_array_to_live_points(live, model.names); live_s["logL"] = model.log_likelihood(live_s)
worst = live_s[np.argmin(live_s["logL"])]
pool = 400_000 * world
prop = bench.build_proposal("c2_realnvp_mlp", model, live_s, local, pool)
for _ in range(3):
    prop.populate(worst, n_samples=pool, max_samples=pool)
    s = prop.samples
    ref = model.log_likelihood(s)
    ok = np.allclose(s["logL"], ref, rtol=1e-12, atol=1e-10)
    print(f"rank {rank}: pool {len(s)} rows, device-sharded logL == host logL: {ok}; max|d| {np.abs(s['logL'] - ref).max():.2e}", flush=True)
    assert ok
dist.destroy_process_group()
