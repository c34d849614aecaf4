import json, os, sys, tempfile, subprocess
import numpy as np
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, os.getcwd())
    from nessai_b200.flowmodel import B200FlowModel
    g = np.load("tests/golden/c2_realnvp_mlp.npz"); cfg = json.loads(str(g["flow_config"]))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    fm = B200FlowModel(flow_config=cfg, output=tempfile.mkdtemp()); fm.initialise()
    fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    zt = torch.randn(1_000_000, 16, device="cuda")
    for _ in range(3): fm.model._inverse(zt)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fm.model._inverse(zt)
    b.record(); torch.cuda.synchronize()
    print(f"{sys.argv[1]:10s} inverse 1e6 rows: {a.elapsed_time(b)/10:.3f} ms")
else:
    for n in ["base", "ng3", "ng2", "noaff", "nomufu", "nosplit", "noall"]:
        env = dict(os.environ, NB200_LIB=os.path.abspath(f"scripts/ubench/lib_{n}.so"))
        subprocess.run([sys.executable, __file__, n], env=env)
