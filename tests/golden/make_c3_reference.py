"""Golden vectors for config C3 (32-D NSF, 6 layers): outputs of the UNMODIFIED reference
flow (``nessai.flows.utils.configure_model`` -> NeuralSplineFlow on the glasflow shim, torch
fp32 on the CPU) and of the float64 oracle for the same randomly initialised weights.

    python tests/golden/make_c3_reference.py

The weights are not stored: they are reproduced from the seed (``FlowSpec.init_state`` is
pinned bit-identical to ``configure_model`` in tests/test_spec.py) plus the perturbation below,
which tests/test_gpu_c3_nsf.py::weights() repeats.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
import oracle.refenv as refenv  # noqa: E402

refenv.activate()
import torch  # noqa: E402
from nessai.flows.utils import configure_model  # noqa: E402

from nessai_b200.spec import FlowSpec  # noqa: E402
from oracle.flow_numpy import NumpyFlow  # noqa: E402

CFG = dict(n_inputs=32, ftype="nsf", n_blocks=6, n_layers=2, n_neurons=64)


def weights(seed=3):
    torch.manual_seed(seed)
    spec = FlowSpec(dict(CFG))
    theta, ints = spec.init_state()
    sd = spec.state_dict_numpy(theta, ints)
    rng = np.random.default_rng(seed)
    for k, v in sd.items():
        if k.endswith("final_layer.weight"):
            sd[k] = (v + 0.3 * rng.standard_normal(v.shape)).astype(np.float32)
        elif k.endswith("final_layer.bias"):
            sd[k] = (v + 0.5 * rng.standard_normal(v.shape)).astype(np.float32)
    return sd


if __name__ == "__main__":
    sd = weights()
    model = configure_model(dict(CFG))
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.eval()
    nf = NumpyFlow(sd, ftype="nsf", net="resnet", hidden_features=64, num_bins=8, tail_bound=5.0)
    z = (np.random.default_rng(1).normal(size=(1024, 32)) * 1.5).astype(np.float32)
    z[:7] *= 4.0  # some coordinates in the linear tails
    with torch.inference_mode():
        x32, lj32 = model.inverse(torch.from_numpy(z))
        zf32, lp32 = model.forward_and_log_prob(x32)
    x64, lj64 = nf.inverse(z.astype(np.float64))
    xin = x32.numpy().astype(np.float64)
    zf64, _ = nf.forward(xin)
    lp64 = nf.log_prob(xin)
    np.savez_compressed(
        os.path.join(HERE, "c3_nsf_reference.npz"), z=z, inv_x=x32.numpy(), inv_logj=lj32.numpy(),
        fwd_z=zf32.numpy(), fwd_logprob=lp32.numpy(), inv_x64=x64, inv_logj64=lj64, fwd_z64=zf64,
        fwd_logprob64=lp64, w_checksum=np.float64(sum(float(np.abs(v).sum()) for v in sd.values())),
    )
    for name, a, b in (("x", x32.numpy(), x64), ("logj", lj32.numpy(), lj64), ("fwd z", zf32.numpy(), zf64),
                       ("logp", lp32.numpy(), lp64)):
        e = np.abs(a - b)
        print(f"reference fp32 {name:6s} vs f64: median {np.median(e):.2e} 99% {np.quantile(e, .99):.2e} max {e.max():.2e}")
