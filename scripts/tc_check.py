"""GPU check of the tcgen05 path against the golden fixture and the generic kernel."""
import json, os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nessai_b200 import _lib
from nessai_b200.flowmodel import B200FlowModel

g = np.load(f"tests/golden/{sys.argv[1] if len(sys.argv) > 1 else "c2_realnvp_mlp"}.npz")
cfg = json.loads(str(g["flow_config"]))
sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
fm.initialise()
fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
lib = _lib.load()
res = {}
for tc in (0, 1):
    lib.nb200_set_tensor_core_path(tc)
    x, lj = fm.inverse(g["z"]); _, lq = fm.sample_and_log_prob(z=g["z"]); z, lp = fm.forward_and_log_prob(g["x"])
    torch.cuda.synchronize()
    res[tc] = (x, lj, lq, z, lp)
    print(f"tc={tc}: inv x {np.abs(x-g['inv_x64']).max():.2e} logj {np.abs(lj-g['inv_logj64']).max():.2e} "
          f"logq rel {np.abs((lq-g['inv_logq'])/g['inv_logq']).max():.2e} | fwd z {np.abs(z-g['fwd_z64']).max():.2e} "
          f"logp rel {np.abs((lp-g['fwd_logprob64'])/g['fwd_logprob64']).max():.2e}", flush=True)
n = 1_000_000
zt = torch.randn(n, 16, device="cuda")
for tc in (0, 1):
    lib.nb200_set_tensor_core_path(tc)
    for _ in range(3): fm.model._inverse(zt)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): out = fm.model._inverse(zt)
    b.record(); torch.cuda.synchronize()
    print(f"tc={tc}: inverse 1e6 rows {a.elapsed_time(b)/5:.3f} ms", flush=True)
    res[("big", tc)] = [o.cpu().numpy() for o in out]
xa, la, qa = res[("big", 0)]; xb, lb, qb = res[("big", 1)]
print("tc vs generic @1e6: x", np.abs(xa-xb).max(), "logj", np.abs(la-lb).max(), "logq", np.abs(qa-qb).max())
