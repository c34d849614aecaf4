"""Dry run of the bodies of tests/test_gpu_zz_tail_accumulate.py on the SIMULATED device
(tests/_simdevice.py) -- checker script, CPU only.  It exercises the test logic itself (index
arithmetic, shapes, tolerances, loop replay) against the oracle-backed C-ABI stand-in, so that a
mistake in a GPU test is found before a GPU is.  Run from the repo root:
    python tests/tools/dryrun_gpu_tests_on_sim.py
"""

import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import _simdevice  # noqa: E402


class SimB200Flow:
    """Stand-in for nessai_b200.flowmodel.B200Flow: keeps the state dict as a NumpyFlow."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.device = torch.device("cpu")
        self._handle = None
        self.spec = types.SimpleNamespace(D=cfg["n_inputs"])

    def load_state_dict(self, sd):
        from oracle.flow_numpy import NumpyFlow

        sd = {k: np.asarray(v) for k, v in sd.items()}
        nf = NumpyFlow(sd, ftype=self.cfg.get("ftype", "realnvp"), net=self.cfg.get("net", "resnet"),
                       activation_name=self.cfg.get("activation", "relu"), hidden_features=self.cfg["n_neurons"])
        self._handle = _simdevice.SimHandle(nf, self.cfg["n_inputs"])

    def eval(self):
        return self

    def _ready(self):
        assert self._handle is not None


class SimB200FlowModel:
    weights_file = None

    def __init__(self, flow_config=None, training_config=None, output=None, rng=None):
        self.flow_config = dict(flow_config)
        self.model = None

    def initialise(self):
        self.model = SimB200Flow(self.flow_config)

    def sample_and_log_prob(self, N=1, z=None):
        x, lq = self.model._handle.flow.sample_and_log_prob(np.asarray(z, dtype=np.float64))
        return x, lq

    def forward_and_log_prob(self, x):
        nf, D = self.model._handle.flow, self.model._handle.D
        z, lj = nf.forward(np.asarray(x, dtype=np.float64))
        return z, -0.5 * np.sum(z * z, axis=1) - 0.5 * D * np.log(2 * np.pi) + lj


def main():
    sim = _simdevice.install()
    import nessai_b200.flowmodel as fm
    import nessai_b200.proposal as proposal

    fm.B200FlowModel = SimB200FlowModel
    proposal.B200FlowProposal._FlowModelClass = SimB200FlowModel
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    for name in ("empty", "zeros", "full", "empty_like"):
        orig = getattr(torch, name)

        def wrapped(*a, _orig=orig, **k):
            if k.get("device") == "cuda":
                k["device"] = "cpu"
            return _orig(*a, **k)

        setattr(torch, name, wrapped)
    import test_gpu_zz_tail_accumulate as t

    tmp = lambda: __import__("pathlib").Path(tempfile.mkdtemp())  # noqa: E731
    runs = [
        ("tail kernel", lambda: [t.test_reparam_tail_kernel_matches_oracle(n, m, p) for n, m, p in
                                 [(5000, None, True), (100_003, -14.0, True), (5000, None, False), (1, None, True)]]),
        ("sum_exp + x64 accept", t.test_sum_exp_and_x64_accept_match_numpy),
        ("general engine", lambda: t.test_general_engine_turn_and_loop_match_oracle(tmp())),
        ("identity kinds", lambda: t.test_identity_kinds_reproduce_the_fused_affine_path(tmp())),
        ("accumulate 1200", lambda: t.test_accumulate_device_loop_matches_oracle(tmp(), 1200, 400)),
        ("accumulate max_samples", lambda: t.test_accumulate_device_loop_matches_oracle(tmp(), 10**6, 2)),
        ("accumulate standalone", lambda: t.test_accumulate_through_the_standalone_proposal(tmp())),
        ("general accumulate", lambda: t.test_general_accumulate_reproduces_affine_accumulate(tmp())),
    ]
    # the EXISTING populate tests (GPU-verified earlier in the round): a regression guard for the
    # shared host code that the new paths hook into (draw_turn / _after_draw / likelihood cut)
    _simdevice.install_fake_cuda_async()
    import test_gpu_populate as tp

    runs += [
        ("existing: fused turn", lambda: [tp.test_fused_turn_matches_numpy_restatement(nm, tmp()) for nm in
                                          ("c2_realnvp_mlp", "c1_realnvp_2d", "d5_realnvp_perm_tanh")]),
        ("existing: populate contract", lambda: tp.test_populate_contract(tmp())),
        ("existing: host prior", lambda: tp.test_host_prior_path_matches_device_prior(tmp())),
        ("existing: min_log_q", lambda: tp.test_min_log_q_truncation(tmp())),
        ("existing: likelihood threshold", lambda: tp.test_likelihood_threshold_truncation_on_device(tmp())),
        ("existing: pool likelihood", lambda: tp.test_pool_likelihood_on_device(tmp())),
    ]
    # __graft_entry__.smoke(): the driver's round-end smoke test (flow parity + one fused populate turn)
    import __graft_entry__ as entry
    import nessai_b200._lib as nlib

    torch.cuda.is_available = lambda: True
    nlib.reset_launch_count = lambda: None
    nlib.launch_count = lambda: len(sim.calls)
    runs.append(("__graft_entry__.smoke", lambda: entry.smoke(train=False)))
    failed = 0
    for name, fn in runs:
        n0 = len(sim.calls)
        try:
            fn()
            print(f"ok      {name}  ({len(sim.calls) - n0} simulated library calls)")
        except Exception as e:  # noqa: BLE001
            failed += 1
            import traceback

            traceback.print_exc()
            print(f"FAILED  {name}: {type(e).__name__}: {e}")
    print(json.dumps({"dryrun_failed": failed, "of": len(runs)}))
    return failed


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
