"""A/B timing of the fused draw kernels under NB200_LIB=<experimental .so>: C2 MLP and ResidualNet
populate draw (1e6 rows), MLP inverse, plus the golden-vector check of the library under test."""
import json, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from nessai_b200.flowmodel import B200FlowModel
from nessai_b200.livepoint import get_dtype
from nessai_b200.proposal import PopulateEngine
label = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("NB200_LIB", "default"))
for name in ("c2_realnvp_mlp", "c2_realnvp_resnet"):
    g = np.load(f"tests/golden/{name}.npz"); cfg = json.loads(str(g["flow_config"]))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    fm = B200FlowModel(flow_config=cfg, training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp()); fm.initialise()
    fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); fm.model.eval()
    x, lq = fm.sample_and_log_prob(z=g["z"])
    err = float(np.abs(lq - g["inv_logq"]).max()), float(np.abs(x - g["inv_x"]).max())
    D, n = 16, 1_000_000
    names = [f"x{i}" for i in range(D)]
    eng = PopulateEngine(fm, names, get_dtype(names))
    eng.configure(np.full(D, 1.5), np.full(D, 0.25), np.full(D, -10.0), np.full(D, 10.0), -D * np.log(20.0), 4.9)
    eng._ensure(n, n, False)
    for _ in range(5): eng.draw_turn(n)
    torch.cuda.synchronize(); ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.draw_turn(n); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"{label:12s} {name:18s} draw 1e6 rows: median {np.median(ts):.4f} ms (min {min(ts):.4f}); golden max|dlogq| {err[0]:.2e} max|dx| {err[1]:.2e}", flush=True)
