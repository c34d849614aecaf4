"""Where does the end-to-end populate() time go? (host-side phases, wall clock with syncs)"""
import os, sys, time, tempfile, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nessai_b200.livepoint import numpy_array_to_live_points
from nessai_b200.proposal import B200FlowProposal, PopulateEngine

g, cfg, sd = bench.load_c2(); live, _ = bench.live_points(); model = bench.GaussianModel()
pool = 1_000_000
prop = B200FlowProposal(model, rng=np.random.default_rng(1), flow_config=cfg, output=tempfile.mkdtemp(), poolsize=pool, drawsize=pool)
prop.initialise()
ls = numpy_array_to_live_points(live, model.names); prop.check_state(ls)
prop.flow.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); prop.flow.model.eval()
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize()
        T[name] = T.get(name, 0) + time.perf_counter() - t; return r
    return w
for _ in range(3): prop.populate(None, n_samples=pool, max_samples=pool)
eng = prop._engine
eng.draw_turn = timed("draw_turn", eng.draw_turn); eng.accept_turn = timed("accept_turn", eng.accept_turn); eng._gather_rows = timed("gather_rows(D2H)", eng._gather_rows)
eng.run = timed("engine.run total", eng.run)
model.log_likelihood = timed("log_likelihood(host)", model.log_likelihood)
N = 10
t0 = time.perf_counter()
for _ in range(N): prop.populate(None, n_samples=pool, max_samples=pool)
tot = time.perf_counter() - t0
for k, v in T.items(): print(f"{k:28s} {1e3*v/N:8.3f} ms / populate")
print(f"{'populate() wall':28s} {1e3*tot/N:8.3f} ms ; population_time {1e3*prop.population_time.total_seconds()/(N+3):.3f} ms; rows out {prop.samples.size}")
