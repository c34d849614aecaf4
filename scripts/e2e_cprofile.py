"""cProfile of B200FlowProposal.populate (host side) at the bench configuration."""
import cProfile, pstats, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nessai_b200.livepoint import numpy_array_to_live_points
from nessai_b200.proposal import B200FlowProposal
g, cfg, sd = bench.load_c2(); live, _ = bench.live_points(); model = bench.GaussianModel()
pool = 1_000_000
prop = B200FlowProposal(model, rng=np.random.default_rng(1), flow_config=cfg, output=tempfile.mkdtemp(), poolsize=pool, drawsize=pool)
prop.initialise()
ls = numpy_array_to_live_points(live, model.names); prop.check_state(ls)
prop.flow.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); prop.flow.model.eval()
eng = prop._get_engine()
import logging; logging.disable(logging.CRITICAL)
for _ in range(3): eng.run(pool, pool, max_samples=pool)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): eng.run(pool, pool, max_samples=pool)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
# the plugin-facing call around it
worst = ls[0]
ls["logL"] = model.log_likelihood(ls)
for _ in range(3): prop.populate(worst, n_samples=pool, max_samples=pool)
import time
stamps = []
orig_get, orig_run = prop._get_engine, eng.run
def get():
    stamps.append(("get0", time.perf_counter())); e = orig_get(); stamps.append(("get1", time.perf_counter())); return e
def run(*a, **k):
    stamps.append(("run0", time.perf_counter())); r = orig_run(*a, **k); stamps.append(("run1", time.perf_counter())); return r
prop._get_engine, eng.run = get, run
orig_ll = model.log_likelihood
def ll(x):
    stamps.append(("ll0", time.perf_counter())); return orig_ll(x)
model.log_likelihood = ll
for rep in range(4):
    torch.cuda.synchronize(); stamps.clear()
    t0 = time.perf_counter(); prop.populate(worst, n_samples=pool, max_samples=pool); t1 = time.perf_counter()
    print("populate total %.3f ms:" % (1e3 * (t1 - t0)), " ".join(f"{n}@{1e3*(t-t0):.3f}" for n, t in stamps))
