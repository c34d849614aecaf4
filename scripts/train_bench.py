"""FlowModel.train on the C2 live points (2000 x 16, the default training_config: batch 1000,
val_size 0.1, lr 1e-3, patience 20, max_epochs 500) -- BASELINE.json north_star's second function.
Prints ONE JSON line.  ``bench.py`` runs it in a subprocess for both arms:

    python scripts/train_bench.py ours        # B200FlowModel.train (persistent training kernel)
    python scripts/train_bench.py reference   # the unmodified reference's FlowModel.train, host cores

Per flow (MLP conditioner = the headline flow, ResidualNet = nessai's default): wall time of
``train()`` (weights file included, plotting off), epochs run, ms per epoch, and for ``ours`` the
device-timed microseconds per optimisation step (CUDA events around cooperative launches of 32
epochs each, no validation, >= 1 s in total).
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
SEED = 20251017
D = 16
FLOWS = {
    "c2_mlp": dict(n_inputs=16, n_neurons=64, n_blocks=4, n_layers=2, ftype="realnvp", net="mlp"),
    "c2_resnet_default": dict(n_inputs=16, n_neurons=64, n_blocks=4, n_layers=2, ftype="realnvp"),
}


def live_points(n=2000):
    rng = np.random.default_rng(SEED)
    idx = np.arange(D)
    cov = 0.5 ** np.abs(idx[:, None] - idx[None, :])
    x = rng.multivariate_normal(np.zeros(D), cov, size=n)
    return (x - x.mean(0)) / x.std(0)  # z-score, the proposal's default reparameterisation


def run_train(FlowModelClass, cfg, x, repeats, **tc):
    import torch

    walls, epochs, hist = [], [], None
    for r in range(repeats + 1):  # the first run is the warm-up
        torch.manual_seed(SEED)
        fm = FlowModelClass(flow_config=dict(cfg), training_config=dict(tc), output=tempfile.mkdtemp(),
                            rng=np.random.default_rng(SEED))
        fm.initialise()
        t0 = time.perf_counter()
        hist = fm.train(x, plot=False)
        if r:
            walls.append(time.perf_counter() - t0)
            epochs.append(len(hist["loss"]))
    return walls, epochs, hist


def main():
    impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
    x = live_points()
    out = {"impl": impl, "data": "C2 live points 2000 x 16 (z-scored), default training_config", "flows": {}}
    if impl == "reference":
        import oracle.refenv as refenv

        refenv.activate()
        import torch
        from nessai.flowmodel import FlowModel

        out["threads"] = torch.get_num_threads()
        for name, cfg in FLOWS.items():
            walls, epochs, hist = run_train(FlowModel, cfg, x, repeats=1)
            out["flows"][name] = {
                "train_wall_s": float(np.median(walls)), "epochs": int(np.median(epochs)),
                "ms_per_epoch": 1e3 * float(np.median(walls)) / float(np.median(epochs)),
                "final_loss": float(hist["loss"][-1]), "final_val_loss": float(hist["val_loss"][-1]),
            }
        print(json.dumps(out), flush=True)
        return
    import torch

    from nessai_b200 import _lib
    from nessai_b200.flowmodel import B200FlowModel

    for name, cfg in FLOWS.items():
        _lib.reset_launch_count()
        walls, epochs, hist = run_train(B200FlowModel, cfg, x, repeats=5, device_tag="cuda:0")
        launches = _lib.launch_count()
        # device-timed optimisation steps: launches of 32 epochs x 2 batches, no validation
        fm = B200FlowModel(flow_config=dict(cfg), training_config=dict(device_tag="cuda:0"), output=tempfile.mkdtemp())
        fm.initialise()
        xt = torch.from_numpy(x[:1800].astype(np.float32)).cuda()
        E, batch = 32, 1000
        perms = torch.stack([torch.randperm(len(xt)) for _ in range(E)]).cuda()
        tr = fm._trainer()
        tr.begin_run()
        e = 0
        for _ in range(2):
            e = tr.run(xt, None, perms, batch, None, None, fm._optimiser, 5.0, [1e-3] * E, e, False, 10**6)[0]
        torch.cuda.synchronize()
        per_step, total = [], 0.0
        while total < 1.0 and len(per_step) < 400:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            e = tr.run(xt, None, perms, batch, None, None, fm._optimiser, 5.0, [1e-3] * E, e, False, 10**6)[0]
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            total += 1e-3 * ms
            per_step.append(1e3 * ms / (E * 2))
        q = np.percentile(per_step, [25, 50, 75])
        out["flows"][name] = {
            "train_wall_s": float(np.median(walls)), "train_wall_s_iqr": float(np.subtract(*np.percentile(walls, [75, 25]))),
            "epochs": int(np.median(epochs)),
            "ms_per_epoch": 1e3 * float(np.median(walls)) / float(np.median(epochs)),
            "final_loss": float(hist["loss"][-1]), "final_val_loss": float(hist["val_loss"][-1]),
            "us_per_step": float(q[1]), "us_per_step_iqr": float(q[2] - q[0]),
            "steps_per_s": 1e6 / float(q[1]), "step_rows": [1000, 800],
            "kernel_launches_per_train": launches / 6.0,
        }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
