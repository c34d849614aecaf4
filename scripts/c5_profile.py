"""cProfile of the per-level flow work of the importance sampler (bench.c5_levels) on the GPU."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nessai_b200.importance import B200ImportanceFlowModel
tc = dict(device_tag="cuda:0")
bench.c5_levels(B200ImportanceFlowModel, 3, 2000, 200_000, tc, dev_sync=torch.cuda.synchronize)  # warm
pr = cProfile.Profile(); pr.enable()
out = bench.c5_levels(B200ImportanceFlowModel, 5, 2000, 200_000, tc, dev_sync=torch.cuda.synchronize)
pr.disable()
print(out)
pstats.Stats(pr).sort_stats("cumtime").print_stats(32)
