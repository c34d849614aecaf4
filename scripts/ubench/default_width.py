"""usage: default_width.py [n_inputs ...]   (default: the whole table)
The reference's DEFAULT flow (ResidualNet conditioner, width 2 * n_inputs, flowmodel/utils.py:39-42) on the
tcgen05 kernels (hidden units zero-padded to 64) against the generic fp32 kernel: inverse + log-prob of 1e6 rows."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from nessai_b200 import _lib
from nessai_b200.flowmodel import B200FlowModel
lib = _lib.load()
n = 1_000_000
only = [int(a) for a in sys.argv[1:]]
for D, net in ((16, "resnet"), (8, "resnet"), (4, "resnet"), (2, "resnet"), (16, "mlp"), (16, "nsf"), (8, "nsf"),
               (20, "resnet"), (24, "resnet"), (32, "resnet")):
    if only and D not in only:
        continue
    cfg = dict(n_inputs=D, n_blocks=4, n_layers=2, ftype="realnvp", net=net)  # n_neurons: the default
    if net == "nsf":
        cfg = dict(n_inputs=D, n_blocks=4, n_layers=2, ftype="nsf")
    torch.manual_seed(D)
    fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
    fm.initialise(); fm.model.eval()
    z = torch.randn(n, D, device="cuda")
    res = {}
    for tc in (1, 0):
        lib.nb200_set_tensor_core_path(tc)
        for _ in range(3): fm.model._inverse(z)
        torch.cuda.synchronize(); ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = fm.model._inverse(z); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res[tc] = (float(np.median(ts)), out[2].float().cpu().numpy())
    lib.nb200_set_tensor_core_path(1)
    err = float(np.nanmax(np.abs(res[1][1] - res[0][1])))
    print(f"D={D:2d} {net:6s} H={fm.flow_config.get('n_neurons', '?')}: tcgen05 {res[1][0]:.3f} ms, generic fp32 {res[0][0]:.3f} ms "
          f"({res[0][0] / res[1][0]:.1f}x) per 1e6 rows; max|dlogq| between the paths {err:.1e}", flush=True)
