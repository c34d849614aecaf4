"""A few epochs of the fused training kernels on a golden fixture (profiling target).
usage: train_prof.py [fixture] [batch] [epochs]"""
import json, os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nessai_b200.flowmodel import B200FlowModel
name = sys.argv[1] if len(sys.argv) > 1 else "c2_realnvp_mlp"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 4
g = np.load(f"tests/golden/{name}.npz")
cfg = json.loads(str(g["flow_config"]))
sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
fm.initialise()
fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
xt = torch.from_numpy(np.asarray(g["train_data"]).astype(np.float32)).cuda()
perm = torch.randperm(len(xt)).cuda()
tr = fm._trainer()
for _ in range(epochs): tr.epoch(xt, None, perm, batch, fm._optimiser, 5.0)
torch.cuda.synchronize()
if len(sys.argv) > 4:  # timing
    n = int(sys.argv[4]); steps = n * ((len(xt) + batch - 1) // batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): tr.epoch(xt, None, perm, batch, fm._optimiser, 5.0)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} batch {batch}: {1e3 * e0.elapsed_time(e1) / steps:.1f} us/step over {steps} steps ({len(xt)} rows)")
