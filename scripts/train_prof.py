import json, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nessai_b200.flowmodel import B200FlowModel
name = sys.argv[1] if len(sys.argv) > 1 else "c2_realnvp_mlp"
g = np.load(f"tests/golden/{name}.npz")
cfg = json.loads(str(g["flow_config"]))
sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
fm.initialise()
fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
xt = torch.from_numpy(np.asarray(g["train_data"]).astype(np.float32)).cuda()
perm = torch.randperm(len(xt)).cuda()
tr = fm._trainer()
for _ in range(4): tr.epoch(xt, None, perm, 1000, fm._optimiser, 5.0)
torch.cuda.synchronize()
