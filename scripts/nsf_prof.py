"""C3 (32-D NSF) sample_and_log_prob of 2e6 rows, a few calls: the ncu target for
flow_tc_nsf_kernel (6 launches per call; capture one with -k regex:flow_tc_nsf -s 8 -c 1)."""
import os, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nessai_b200.flowmodel import B200FlowModel

torch.manual_seed(3)
fm = B200FlowModel(flow_config=dict(n_inputs=32, ftype="nsf", n_blocks=6, n_layers=2, n_neurons=64),
                   training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
fm.initialise()
fm.model.eval()
z = torch.randn(2_000_000, 32, device="cuda")
for _ in range(3):
    fm.model._inverse(z)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fm.model._inverse(z)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("c3 inverse 2e6 rows ms:", " ".join(f"{t:.3f}" for t in ts))
