"""GPU check of the fused training kernels against the float64 oracle (one case, verbose)."""
import faulthandler; faulthandler.dump_traceback_later(40, exit=True)
import json, os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root
from nessai_b200.flowmodel import B200FlowModel
from oracle.train_numpy import TrainStepOracle

name = sys.argv[1] if len(sys.argv) > 1 else "c2_realnvp_mlp"
n_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 301
g = np.load(f"tests/golden/{name}.npz")
cfg = json.loads(str(g["flow_config"]))
sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
fm.initialise()
fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
spec = fm.model.spec
x = np.asarray(g["train_data"])[:n_rows].astype(np.float32)
theta64 = fm.model.theta_numpy().astype(np.float64)
loss64, grad64 = TrainStepOracle(spec, fm.model.ints).loss_and_grad(theta64, x.astype(np.float64))
loss, grad, info = fm._trainer().loss_and_grad(torch.from_numpy(x).cuda())
torch.cuda.synchronize()
grad = grad.cpu().numpy().astype(np.float64)
print(f"{name} n={n_rows}: loss {float(loss):.6f} vs {loss64:.6f}; |g| {float(info[1]):.5f} vs {np.linalg.norm(grad64):.5f}")
for e in spec.entries:
    if e.kind != "param": continue
    a, b = grad[e.offset:e.offset + e.size], grad64[e.offset:e.offset + e.size]
    r = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
    flag = "" if r < 2e-4 else "   <-- MISMATCH"
    print(f"  {e.key[-60:]:60s} |b|={np.linalg.norm(b):.3e} rel={r:.2e}{flag}")
if len(sys.argv) > 3:
    # timing of whole epochs
    xt = torch.from_numpy(np.asarray(g["train_data"]).astype(np.float32)).cuda()
    perm = torch.randperm(len(xt)).cuda()
    tr = fm._trainer()
    for _ in range(3): tr.epoch(xt, None, perm, 1000, fm._optimiser, 5.0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n_ep = 50
    for _ in range(n_ep): tr.epoch(xt, None, perm, 1000, fm._optimiser, 5.0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nb = -(-len(xt) // 1000)
    print(f"epoch of {len(xt)} rows ({nb} batches): {1e3*dt/n_ep:.3f} ms  -> {1e6*dt/n_ep/nb:.1f} us/step")
    # comparison: torch autograd (tests/_eager_flow.py) + torch.optim.AdamW on the same GPU
    sys.path.insert(0, "tests")
    from _eager_flow import EagerFlow
    eager = EagerFlow(spec, fm.model.ints, fm.model.device)
    tp = torch.nn.Parameter(fm.model.theta_p.detach().clone()); tb = fm.model.theta_b.clone()
    opt = torch.optim.AdamW([tp], lr=1e-3)
    def step(xb):
        opt.zero_grad(set_to_none=True)
        loss = -eager.log_prob((tp, tb), xb, training=True).mean()
        loss.backward(); torch.nn.utils.clip_grad_norm_([tp], 5.0); opt.step()
    for _ in range(3): step(xt[:1000])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): step(xt[:1000])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"torch autograd + AdamW on the same GPU: {1e6*dt/20:.1f} us/step")
