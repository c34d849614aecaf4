"""Where the host time of B200FlowModel.train goes (cProfile of warm calls on the C2 live points)."""
import cProfile, io, os, pstats, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.train_bench import FLOWS, SEED, live_points
from nessai_b200.flowmodel import B200FlowModel
name = sys.argv[1] if len(sys.argv) > 1 else "c2_resnet_default"
x = live_points()
def make():
    torch.manual_seed(SEED)
    fm = B200FlowModel(flow_config=dict(FLOWS[name]), training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp(), rng=np.random.default_rng(SEED))
    fm.initialise()
    return fm
fm = make(); fm.train(x, plot=False)  # warm-up
ts = []
pr = cProfile.Profile()
for r in range(5):
    fm = make()
    t0 = time.perf_counter(); pr.enable(); h = fm.train(x, plot=False); pr.disable(); ts.append(time.perf_counter() - t0)
print(name, "train() wall ms:", [round(1e3 * t, 1) for t in ts], "epochs", len(h["loss"]))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
