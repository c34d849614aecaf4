"""Optimisation steps of the fused training kernels on a golden fixture (profiling / timing target).
usage: train_prof.py [fixture] [batch] [warm-up runs] [timed runs]
Each run = one cooperative launch of 32 epochs over the fixture's 2000 training rows (no validation,
no early stopping): us/step = CUDA-event time / optimisation steps."""
import json, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nessai_b200.flowmodel import B200FlowModel
name = sys.argv[1] if len(sys.argv) > 1 else "c2_realnvp_mlp"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
timed = int(sys.argv[4]) if len(sys.argv) > 4 else 0
g = np.load(f"tests/golden/{name}.npz")
cfg = json.loads(str(g["flow_config"]))
sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
fm.initialise()
fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
xt = torch.from_numpy(np.asarray(g["train_data"]).astype(np.float32)).cuda()
E = 32
perms = torch.stack([torch.randperm(len(xt)) for _ in range(E)]).cuda()
tr = fm._trainer()
tr.begin_run()
run = lambda e0: tr.run(xt, None, perms, batch, None, None, fm._optimiser, 5.0, [1e-3] * E, e0, False, 10**6)
e = 0
for _ in range(warm):
    e = run(e)[0]
torch.cuda.synchronize()
if timed:
    steps = timed * E * ((len(xt) + batch - 1) // batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(timed):
        e = run(e)[0]
    e1.record(); torch.cuda.synchronize()
    print(f"{name} batch {batch}: {1e3 * e0.elapsed_time(e1) / steps:.1f} us/step over {steps} steps ({len(xt)} rows)")
