"""Host timeline of B200FlowProposal.populate (NB200_TRACE): where the end-to-end time goes."""
import os, sys, time, tempfile, datetime
os.environ["NB200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from nessai_b200.livepoint import numpy_array_to_live_points
from nessai_b200.proposal import B200FlowProposal

g, cfg, sd = bench.load_c2()
live, _ = bench.live_points()
model = bench.GaussianModel()
pool = 1_000_000
prop = B200FlowProposal(model, rng=np.random.default_rng(1), flow_config=cfg, training_config=dict(device_tag="cuda:0"),
                        output=tempfile.mkdtemp(), poolsize=pool, drawsize=pool, device_prior="auto")
prop.initialise()
live_s = numpy_array_to_live_points(live, model.names)
live_s["logL"] = model.log_likelihood(live_s)
prop.check_state(live_s)
prop.flow.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
prop.flow.model.eval()
worst = live_s[np.argmin(live_s["logL"])]
for _ in range(3):
    prop.populate(worst, n_samples=pool, max_samples=pool)
for rep in range(3):
    torch.cuda.synchronize()
    prop.population_time *= 0
    t0 = time.perf_counter()
    prop.populate(worst, n_samples=pool, max_samples=pool)
    tr = prop._engine.last_trace
    print(f"populate: population_time {prop.population_time.total_seconds()*1e3:.3f} ms; run() entered at +{(tr[0][1]-t0)*1e3:.3f} ms")
    for (a, ta), (b, tb) in zip(tr[:-1], tr[1:]):
        print(f"   {a:>14s} -> {b:<14s} {1e3*(tb-ta):7.3f} ms")
