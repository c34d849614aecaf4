"""C3 (reference-trained 32-D NSF, tests/golden/c3_nsf_trained.npz) sample_and_log_prob of 2e6 rows, a few
calls: the ncu target for flow_tc_nsf_kernel (6 launches per call; capture one with -k regex:flow_tc_nsf -s 8 -c 1)."""
import json, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nessai_b200.flowmodel import B200FlowModel
g = np.load("tests/golden/c3_nsf_trained.npz"); cfg = json.loads(str(g["flow_config"]))
sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
fm.initialise()
fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); fm.model.eval()
z = torch.randn(2_000_000, 32, device="cuda")
for _ in range(3):
    fm.model._inverse(z)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fm.model._inverse(z); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("c3 inverse 2e6 rows ms:", " ".join(f"{t:.3f}" for t in ts))
