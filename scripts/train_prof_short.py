"""One short cooperative training launch (8 epochs x 2 steps, C2 MLP): the ncu target for tr_train_kernel."""
import json, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nessai_b200.flowmodel import B200FlowModel
g = np.load("tests/golden/c2_realnvp_mlp.npz"); cfg = json.loads(str(g["flow_config"]))
fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp()); fm.initialise()
xt = torch.from_numpy(np.asarray(g["train_data"]).astype(np.float32)).cuda()
E = 8
perms = torch.stack([torch.randperm(len(xt)) for _ in range(E)]).cuda()
tr = fm._trainer(); tr.begin_run(); e = 0
for _ in range(3):
    e = tr.run(xt, None, perms, 1000, None, None, fm._optimiser, 5.0, [1e-3] * E, e, False, 10**6)[0]
torch.cuda.synchronize()
