"""Generate the golden fixtures in this directory.

Runs the UNMODIFIED reference (``baseline/_ref``: nessai's FlowModel /
configure_model) on top of ``oracle/shims`` (the restated glasflow.nflows) in
THIS container and freezes weights + input/output vectors as small ``.npz``
files.  /root/reference and the shim cannot be imported by the product, and the
reference install does not have to exist on the GPU box for the ``-m gpu``
tests: they only read the committed ``.npz`` files.

    python tests/golden/make_golden.py

Each fixture holds: the reference ``state_dict`` (keys prefixed ``sd/``), the
``flow_config`` (json), inputs ``x`` / ``z`` and the reference's fp32 outputs
``fwd_z, fwd_logj, fwd_logprob`` (``forward_and_log_prob``) and
``inv_x, inv_logj, inv_logq`` (``inverse`` and ``sample_and_log_prob(z=z)``),
plus float64 values from ``oracle/flow_numpy.py`` (suffix ``64``).
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

import oracle.refenv as refenv  # noqa: E402

refenv.activate()

import torch  # noqa: E402
from nessai.flowmodel import FlowModel  # noqa: E402

from oracle.flow_numpy import NumpyFlow  # noqa: E402

SEED = 20251017


def gaussian_live_points(n, d, rng):
    """Config C2 live points: N(0, Sigma), Sigma_ij = 0.5^|i-j|."""
    idx = np.arange(d)
    cov = 0.5 ** np.abs(idx[:, None] - idx[None, :])
    return rng.multivariate_normal(np.zeros(d), cov, size=n)


def rosenbrock_like(n, d, rng):
    """Curved, non-Gaussian training set on [-5, 5]^d for the spline flow."""
    x = np.empty((n, d))
    x[:, 0] = rng.normal(1.0, 0.5, n)
    for i in range(1, d):
        x[:, i] = 0.5 * x[:, i - 1] ** 2 + rng.normal(0, 0.3, n)
    return np.clip(x, -4.9, 4.9)


def rosenbrock_chain(n, d, rng):
    """Curved, non-Gaussian live points for any d (the quadratic recursion of rosenbrock_like
    diverges beyond a few dimensions): x_i = 0.25 x_{i-1}^2 + 0.75 + noise contracts to the
    Rosenbrock ridge x_{i+1} ~ x_i^2 around (1, ..., 1); inside [-5, 5]^d."""
    x = np.empty((n, d))
    x[:, 0] = rng.normal(1.0, 0.3, n)
    for i in range(1, d):
        x[:, i] = 0.25 * x[:, i - 1] ** 2 + 0.75 + rng.normal(0, 0.1, n)
    return np.clip(x, -4.9, 4.9)


def make(name, flow_config, data, n_eval, epochs, tmp):
    rng = np.random.default_rng(SEED)
    torch.manual_seed(SEED)
    fm = FlowModel(
        flow_config=dict(flow_config),
        training_config=dict(max_epochs=epochs, patience=epochs, batch_size=1000),
        output=os.path.join(tmp, name),
        rng=rng,
    )
    fm.initialise()
    init_sd = {k: v.clone().numpy() for k, v in fm.model.state_dict().items()}
    mu, sd = data.mean(0), data.std(0)
    xp = (data - mu) / sd  # z-score, as the proposal's default reparameterisation
    hist = fm.train(xp, plot=False)
    model = fm.model
    model.eval()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(SEED + 1)
    base_var = float((flow_config.get("distribution_kwargs") or {}).get("var", 1.0))
    z = torch.randn(n_eval, xp.shape[1], generator=g)
    if base_var != 1.0:
        z = z * float(np.sqrt(base_var))
    x = torch.from_numpy(xp[:n_eval]).float()
    with torch.inference_mode():
        fwd_z, fwd_lp = model.forward_and_log_prob(x)
        _, fwd_lj = model.forward(x)
        inv_x, inv_lj = model.inverse(z)
        inv_lq = model.base_distribution_log_prob(z) - inv_lj
    kw = dict(
        ftype={"nsf": "nsf", "maf": "maf"}.get(str(flow_config.get("ftype")).lower(), "realnvp"),
        net=flow_config.get("net", "resnet"),
        activation_name=flow_config.get("activation", "relu"),
        volume_preserving=flow_config.get("use_volume_preserving", False),
        num_bins=flow_config.get("num_bins", 8),
        tail_bound=flow_config.get("tail_bound", 5.0),
        hidden_features=flow_config["n_neurons"],
        base_var=base_var,
    )
    nf = NumpyFlow(state, **kw)
    z64, lj64 = nf.forward(x.numpy())
    lp64 = nf.base_log_prob(z64) + lj64
    x64, ilj64 = nf.inverse(z.numpy())
    out = {f"sd/{k}": v.numpy() for k, v in state.items()}
    out.update({f"init/{k}": v for k, v in init_sd.items()})
    out.update(
        flow_config=np.array(json.dumps(flow_config)),
        x=x.numpy(),
        z=z.numpy(),
        fwd_z=fwd_z.numpy(),
        fwd_logj=fwd_lj.numpy(),
        fwd_logprob=fwd_lp.numpy(),
        inv_x=inv_x.numpy(),
        inv_logj=inv_lj.numpy(),
        inv_logq=inv_lq.numpy(),
        fwd_z64=z64,
        fwd_logj64=lj64,
        fwd_logprob64=lp64,
        inv_x64=x64,
        inv_logj64=ilj64,
        train_data=xp.astype(np.float64),
        loss=np.array(hist["loss"]),
        val_loss=np.array(hist["val_loss"]),
    )
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    err = np.abs(fwd_lp.numpy() - lp64).max()
    print(
        f"{name}: epochs={len(hist['loss'])} loss={hist['loss'][-1]:.3f} "
        f"val={hist['val_loss'][-1]:.3f} |shim-f64| logprob={err:.2e} "
        f"-> {os.path.getsize(path) / 1024:.0f} KiB"
    )


def main():
    import tempfile

    tmp = tempfile.mkdtemp()
    only = set(sys.argv[1:])  # e.g. `make_golden.py d8_maf`: regenerate just that fixture
    global make
    _make = make

    def make(name, *a, **k):  # noqa: F811
        if not only or name in only:
            _make(name, *a, **k)

    rng = np.random.default_rng(SEED)
    live16 = gaussian_live_points(2000, 16, rng)
    # C2: 16-D RealNVP, 4 coupling layers, [64, 64] MLP conditioner
    make(
        "c2_realnvp_mlp",
        dict(n_inputs=16, n_neurons=64, n_blocks=4, n_layers=2, ftype="realnvp", net="mlp"),
        live16, 512, 200, tmp,
    )
    # C2': default ResidualNet conditioner
    make(
        "c2_realnvp_resnet",
        dict(n_inputs=16, n_neurons=64, n_blocks=4, n_layers=2, ftype="realnvp"),
        live16, 512, 100, tmp,
    )
    # small odd-sized variants (ragged dims, other activations / linear transforms)
    live5 = gaussian_live_points(1000, 5, rng)
    make(
        "d5_realnvp_perm_tanh",
        dict(n_inputs=5, n_neurons=10, n_blocks=3, n_layers=1, ftype="realnvp",
             net="mlp", linear_transform="permutation", activation="tanh"),
        live5, 257, 30, tmp,
    )
    make(
        "d4_realnvp_additive_silu",
        dict(n_inputs=4, n_neurons=8, n_blocks=2, n_layers=2, ftype="realnvp",
             linear_transform=None, batch_norm_between_layers=False,
             use_volume_preserving=True, activation="swish"),
        live5[:, :4], 100, 30, tmp,
    )
    # MultivariateNormal base distribution, N(0, 2.5 I) (flows/distributions.py:17-73)
    make(
        "d5_realnvp_mvn",
        dict(n_inputs=5, n_neurons=10, n_blocks=3, n_layers=1, ftype="realnvp", net="mlp",
             distribution="mvn", distribution_kwargs=dict(var=2.5)),
        live5, 257, 30, tmp,
    )
    # config 1: 2-D, 2 coupling layers, all defaults (resnet, LU, BN, n_neurons=4)
    make(
        "c1_realnvp_2d",
        dict(n_inputs=2, n_neurons=4, n_blocks=2, n_layers=2, ftype="realnvp"),
        gaussian_live_points(500, 2, rng), 128, 50, tmp,
    )
    # C3-like spline flow (small D so the fixture stays small) and full C3 shape
    make(
        "d6_nsf",
        dict(n_inputs=6, n_neurons=16, n_blocks=3, n_layers=2, ftype="nsf"),
        rosenbrock_like(2000, 6, rng), 512, 60, tmp,
    )
    # C3 at the configuration's own size (BASELINE.json configs[2]): 32-D neural-spline flow, 6 coupling
    # layers, ResidualNet(64), trained by the reference on a curved 32-D live-point set
    make(
        "c3_nsf_trained",
        dict(n_inputs=32, n_neurons=64, n_blocks=6, n_layers=2, ftype="nsf"),
        rosenbrock_chain(2000, 32, np.random.default_rng(SEED + 32)), 512, 80, tmp,
    )
    # masked autoregressive flow: MADE with residual blocks, reverse permutations
    make(
        "d8_maf",
        dict(n_inputs=8, n_neurons=32, n_blocks=3, n_layers=2, ftype="maf"),
        rosenbrock_like(2000, 8, np.random.default_rng(SEED + 8)), 512, 60, tmp,
    )


if __name__ == "__main__":
    main()
