"""Verbose error breakdown of the 32-D spline-flow kernels against the float64 oracle (checker script;
run from the repo root: python tests/tools/c3_diag.py)."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from test_gpu_c3_nsf import make
from oracle.flow_numpy import NumpyFlow
fm, sd = make(tempfile.mkdtemp())
nf = NumpyFlow(sd, ftype="nsf", net="resnet", hidden_features=64, num_bins=8, tail_bound=5.0)
sd32 = {k: (v.astype(np.float32) if v.dtype.kind == "f" else v) for k, v in sd.items()}
rng = np.random.default_rng(1)
z = (rng.normal(size=(4096, 32)) * 1.5).astype(np.float32)
x, logj = fm.inverse(z.astype(np.float64))
x64, lj64 = nf.inverse(z.astype(np.float64))
def q(name, a, b):
    e = np.abs(a - b)
    print(f"{name:28s} median {np.median(e):.2e}  99% {np.quantile(e, .99):.2e}  99.9% {np.quantile(e, .999):.2e}  max {e.max():.2e}")
q("kernel x vs f64", x, x64); q("kernel logj vs f64", logj, lj64)
zf, logp = fm.forward_and_log_prob(x64)
zf64, lf64 = nf.forward(x64.astype(np.float32).astype(np.float64))
q("kernel fwd z vs f64", zf, zf64); q("kernel fwd logp vs f64", logp, nf.log_prob(x64.astype(np.float32).astype(np.float64)))
# the same arithmetic in float32 numpy (what an fp32 CPU reference does)
try:
    nf32 = NumpyFlow(sd32, ftype="nsf", net="resnet", hidden_features=64, num_bins=8, tail_bound=5.0)
    x32, lj32 = nf32.inverse(z)
    print("numpy dtype", x32.dtype)
    q("fp32 numpy x vs f64", x32, x64); q("fp32 numpy logj vs f64", lj32, lj64)
except Exception as e:
    print("fp32 numpy failed", e)
