"""Phase timeline of one epilogue group / issuer of the tcgen05 kernel (NB200_TC_TRACE build)."""
import ctypes as C, json, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.getcwd())
os.environ["NB200_LIB"] = os.path.abspath("scripts/ubench/lib_trace.so")
from nessai_b200 import _lib
from nessai_b200.flowmodel import B200FlowModel
g = np.load("tests/golden/c2_realnvp_mlp.npz"); cfg = json.loads(str(g["flow_config"]))
sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
fm = B200FlowModel(flow_config=cfg, output=tempfile.mkdtemp()); fm.initialise()
fm.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
zt = torch.randn(1_000_000, 16, device="cuda")
lib = C.CDLL(os.environ["NB200_LIB"])
buf = (C.c_longlong * 8192)(); n = (C.c_int * 2)()
for _ in range(3):
    fm.model._inverse(zt); lib.nb200_debug_trace(buf, n)
fm.model._inverse(zt); lib.nb200_debug_trace(buf, n)
a = np.frombuffer(buf, dtype=np.int64).reshape(2, 4096)
ev = []
for who in (0, 1):
    for v in a[who, : n[who]]:
        ev.append((int(v) >> 8, who, int(v) & 255))
ev.sort()
t0 = ev[0][0]
names = {7: "E3 TMEM loaded", 8: "E3 coupling done", 9: "E0 split done", 1: "E0 done->arrive", 2: "E1 wake", 3: "E1 done->arrive", 4: "E2 wake", 5: "E2 done->arrive", 6: "E3 wake",
         11: "  I: wake G1", 12: "  I: G1 issued+commit", 13: "  I: wake G2", 14: "  I: G2 issued+commit", 15: "  I: wake G3", 16: "  I: G3 issued+commit"}
prev = t0
for t, who, tag in ev[60:110]:
    print(f"{t - t0:9d} (+{t - prev:5d})  {names[tag]}")
    prev = t
# per-phase statistics over the whole run (epilogue side)
e = [(t, tag) for t, who, tag in ev if who == 0]
d = {}
for (t1, g1), (t2, g2) in zip(e[:-1], e[1:]):
    d.setdefault((g1, g2), []).append(t2 - t1)
for k, v in sorted(d.items()):
    print(f"epilogue {names[k[0]]:18s} -> {names[k[1]]:18s}: median {int(np.median(v)):6d} cycles (n={len(v)})")
i = [(t, tag) for t, who, tag in ev if who == 1]
d = {}
for (t1, g1), (t2, g2) in zip(i[:-1], i[1:]):
    d.setdefault((g1, g2), []).append(t2 - t1)
for k, v in sorted(d.items()):
    print(f"issuer {names[k[0]]:24s} -> {names[k[1]]:24s}: median {int(np.median(v)):6d} cycles (n={len(v)})")
