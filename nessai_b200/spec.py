"""Flow description, flat parameter layout and eval-mode folding (host side).

The reference assembles its flows from ``glasflow.nflows`` modules
(/root/reference/src/nessai/flows/realnvp.py:76-214, nsf.py:60-130) and keeps
the parameters in a ``torch.nn.Module`` tree.  Here a flow is

* a :class:`FlowSpec` -- the architecture derived from the *same*
  ``flow_config`` dictionary (``configure_model``,
  /root/reference/src/nessai/flows/utils.py:208-246);
* ONE flat fp32 buffer ``theta`` holding every trainable parameter followed by
  the float buffers (BatchNorm running statistics), laid out in the order of the
  reference ``state_dict`` so weight files interchange (SURVEY.md section 8c,
  "State-dict key layout"; pinned against the reference on the restated nflows layer of
  ``oracle/shims`` -- real glasflow is absent from this image, see
  ``tests/test_spec.py::test_interchange_with_real_glasflow``);
* a *program*: the eval-mode flow folded on the host (float64) into the op
  list the CUDA kernels interpret (``csrc/flow_program.h``).  Every
  row-constant transform (permutation, LU, eval-mode BatchNorm) between two
  coupling transforms is folded into one dense DxD affine, and the coupling's
  gather/scatter by mask is folded into that affine as a row permutation so the
  kernels always see ``[identity | transformed]`` halves.

Nothing here touches CUDA; it is plain numpy/torch-CPU plumbing.
"""

from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# --- op codes shared with csrc/flow_program.h ---------------------------------
OP_LINEAR = 0
OP_COUPLING_AFFINE = 1
OP_COUPLING_SPLINE = 2

FLAG_IN_ACT = 1
FLAG_OUT_ACT = 2
FLAG_ACCUM = 4
FLAG_INVERSE = 8
FLAG_ADDITIVE = 16
FLAG_SOFTPLUS_SCALE = 32  # scale = softplus(u) + 1e-3 (masked affine autoregressive transform)
FLAG_NO_LOGDET = 64  # do not accumulate log|det| (intermediate passes of the MAF inverse)

BUF_X0, BUF_X1, BUF_A0, BUF_A1 = 0, 1, 2, 3

ACT_RELU, ACT_TANH, ACT_SILU = 0, 1, 2
_ACTIVATIONS = {"relu": ACT_RELU, "tanh": ACT_TANH, "swish": ACT_SILU, "silu": ACT_SILU}

OP_INTS = 16  # ints per encoded op


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def get_n_neurons(n_neurons=None, n_inputs=None, default=8) -> int:
    """Mirror of /root/reference/src/nessai/flows/utils.py:105-165."""
    if n_inputs is None:
        n = default if n_neurons is None else n_neurons
    else:
        if n_neurons is None or n_neurons in {"auto", "double"}:
            n = 2 * n_inputs
        elif n_neurons == "equal":
            n = n_inputs
        elif n_neurons == "half":
            n = n_inputs // 2
        else:
            n = n_neurons
    try:
        return int(n)
    except ValueError:
        raise ValueError(
            "Could not get number of neurons. `n_neurons` was set to "
            f"`{n_neurons}` which could not be translated to a valid int."
        )


@dataclass
class Entry:
    """One tensor of the reference ``state_dict``."""

    key: str
    shape: Tuple[int, ...]
    kind: str  # "param" | "fbuf" (float buffer) | "ibuf" (int64 buffer)
    offset: int = -1  # float offset into theta (param / fbuf)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if len(self.shape) else 1


@dataclass
class LinearRef:
    """Keys of one ``nn.Linear`` (weight (out,in), bias (out,))."""

    weight: str
    bias: str
    n_in: int
    n_out: int
    mask: Optional[str] = None  # MADE: float buffer multiplied into the weight


@dataclass
class LayerSpec:
    """One coupling layer and the row-constant transforms around it."""

    perm_key: Optional[str] = None
    lu_prefix: Optional[str] = None  # keys: bias, lower_entries, upper_entries, unconstrained_upper_diag
    coupling_prefix: str = ""
    identity: np.ndarray = None  # int64 feature indices
    transform: np.ndarray = None
    linears: List[LinearRef] = field(default_factory=list)
    bn_prefix: Optional[str] = None


class FlowSpec:
    """Architecture + flat layout for the flows nessai builds.

    Accepts the reference's ``flow_config`` keys
    (/root/reference/src/nessai/flowmodel/config.py:12-24 plus the RealNVP / NSF
    keyword arguments).  Unsupported options raise ``NotImplementedError``
    loudly -- there is no CPU fallback.
    """

    LU_EPS = 1e-3
    BN_EPS = 1e-5
    BN_MOMENTUM = 0.1

    def __init__(self, flow_config: dict):
        cfg = copy.deepcopy(dict(flow_config))
        cfg.pop("inference_device_tag", None)
        if cfg.pop("flow", None) is not None:
            raise NotImplementedError(
                "nessai_b200: custom flow classes (flow_config['flow']) are not "
                "supported; use ftype='realnvp' or 'nsf'"
            )
        # base distribution (/root/reference/src/nessai/flows/utils.py:35-102): None is nflows'
        # StandardNormal; "mvn" / "normal" is nessai's MultivariateNormal, N(0, var I)
        # (flows/distributions.py:17-73), var from distribution_kwargs
        dist = cfg.pop("distribution", None)
        dist_kwargs = dict(cfg.pop("distribution_kwargs", None) or {})
        self.base_var = 1.0
        if dist is not None:
            name = dist.lower() if isinstance(dist, str) else getattr(dist, "__name__", str(dist))
            if name in ("mvn", "normal", "MultivariateNormal"):
                self.base_var = float(dist_kwargs.pop("var", 1))
                if not (self.base_var > 0.0 and np.isfinite(self.base_var)):
                    raise ValueError(f"MultivariateNormal needs a positive finite variance, got {self.base_var}")
                if dist_kwargs:
                    raise TypeError(f"MultivariateNormal got unexpected arguments: {sorted(dist_kwargs)}")
            elif isinstance(dist, str) and name not in ("lars", "resampled", "uniform"):
                raise ValueError(f"Unknown distribution: {dist}")
            else:
                raise NotImplementedError(
                    f"nessai_b200: base distribution {dist!r} is not implemented "
                    "(StandardNormal and MultivariateNormal are)"
                )
        n_inputs = cfg.pop("n_inputs")
        if not isinstance(n_inputs, (int, np.integer)) or isinstance(n_inputs, bool):
            raise TypeError("Number of inputs (n_inputs) must be an int")
        self.D = int(n_inputs)
        if self.D <= 1:
            raise ValueError(
                f"RealNVP requires at least 2 dimensions. Specified dimensions: {self.D}."
            )
        self.H = get_n_neurons(cfg.pop("n_neurons", None), self.D)
        self.L = int(cfg.pop("n_blocks", 4))  # number of coupling transforms
        self.n_layers = int(cfg.pop("n_layers", 2))  # conditioner depth
        ftype = str(cfg.pop("ftype", "realnvp")).lower()
        if ftype in ("realnvp", "frealnvp"):
            self.ftype = "realnvp"
        elif ftype in ("nsf", "spline"):
            self.ftype = "nsf"
        elif ftype == "maf":
            self.ftype = "maf"
        else:
            raise NotImplementedError(
                f"nessai_b200: flow type {ftype!r} is not implemented (realnvp, nsf, maf)"
            )
        act = cfg.pop("activation", "relu")
        if callable(act):
            act = getattr(act, "__name__", str(act))
        if act not in _ACTIVATIONS:
            raise ValueError(f"Unknown activation: {act}")
        self.activation_name = act
        self.activation = _ACTIVATIONS[act]

        for key, allowed in (
            ("context_features", (None,)),
            ("dropout_probability", (0.0, 0, None)),
            ("batch_norm_within_layers", (False, None)),
            ("pre_transform", (None,)),
            ("actnorm", (False, None)),
        ):
            if cfg.get(key, allowed[0]) not in allowed:
                raise NotImplementedError(
                    f"nessai_b200: flow_config[{key!r}]={cfg[key]!r} is not implemented"
                )
            cfg.pop(key, None)
        cfg.pop("pre_transform_kwargs", None)

        if self.ftype == "realnvp":
            self.net = str(cfg.pop("net", "resnet")).lower()
            if self.net not in ("resnet", "mlp"):
                raise ValueError(
                    f"Unknown nn type: {self.net}. Choose from: {{resnet, mlp}}."
                )
            self.volume_preserving = bool(cfg.pop("use_volume_preserving", False))
            self.bn_between = bool(cfg.pop("batch_norm_between_layers", True))
            self.linear_transform = cfg.pop("linear_transform", "lu")
            mask = cfg.pop("mask", None)
            self.num_bins = 0
            self.tail_bound = 0.0
        elif self.ftype == "maf":
            # /root/reference/src/nessai/flows/maf.py:62-104
            self.net = "resnet" if cfg.pop("use_residual_blocks", True) else "mlp"
            if cfg.pop("use_random_masks", False):
                raise NotImplementedError("nessai_b200: MAF with random masks is not implemented")
            self.random_permutations = bool(cfg.pop("use_random_permutations", False))
            self.volume_preserving = False
            self.bn_between = bool(cfg.pop("batch_norm_between_layers", False))
            self.linear_transform = "permutation"
            self.num_bins = 0
            self.tail_bound = 0.0
            mask = None
        else:
            self.net = "resnet"
            self.volume_preserving = False
            self.bn_between = bool(cfg.pop("batch_norm_between_layers", False))
            self.linear_transform = cfg.pop("linear_transform", "permutation")
            self.num_bins = int(cfg.pop("num_bins", 8))
            if cfg.pop("tails", "linear") != "linear":
                raise NotImplementedError("nessai_b200: only tails='linear' is implemented")
            self.tail_bound = float(cfg.pop("tail_bound", 5.0))
            if cfg.pop("apply_unconditional_transform", False):
                raise NotImplementedError(
                    "nessai_b200: apply_unconditional_transform is not implemented"
                )
            mask = None
        if isinstance(self.linear_transform, str):
            self.linear_transform = self.linear_transform.lower()
            if self.linear_transform == "none":
                self.linear_transform = None
        if self.linear_transform not in (None, "lu", "permutation"):
            if self.linear_transform == "svd":
                raise NotImplementedError("nessai_b200: linear_transform='svd' is not implemented")
            raise ValueError(
                f"Unknown linear transform: {self.linear_transform}. "
                "Choose from: {permutation, lu, svd}."
            )
        if cfg:
            raise NotImplementedError(
                f"nessai_b200: unsupported flow_config keys: {sorted(cfg)}"
            )
        self.masks = self._make_masks(mask)
        self._build_layout()

    # ------------------------------------------------------------------ masks
    def _make_masks(self, mask) -> np.ndarray:
        D, L = self.D, self.L
        if self.ftype == "maf":
            return np.ones((L, D))  # every feature is transformed (autoregressively)
        if self.ftype == "nsf":
            # create_alternating_binary_mask(features, even=(i % 2 == 0)):
            # ones (= transformed) at start::2, start = 0 if even else 1
            out = np.zeros((L, D))
            for i in range(L):
                out[i, (0 if i % 2 == 0 else 1) :: 2] = 1.0
            return out
        # /root/reference/src/nessai/flows/realnvp.py:114-131
        if mask is None:
            m = np.ones(D)
            m[::2] = -1
        else:
            m = np.array(mask, dtype=np.float64)
            if not m.shape[-1] == D:
                raise ValueError("Mask does not match number of features")
            if m.ndim == 2 and not m.shape[0] == L:
                raise ValueError("Mask does not match number of layers")
        if m.ndim == 1:
            out = np.empty((L, D))
            m = m.copy()
            for i in range(L):
                out[i] = m
                m *= -1
            return out
        return m

    # ----------------------------------------------------------------- layout
    @property
    def coupling_multiplier(self) -> int:
        if self.ftype == "nsf":
            return 3 * self.num_bins - 1
        return 1 if (self.volume_preserving and self.ftype == "realnvp") else 2

    def _build_layout(self) -> None:
        D, H = self.D, self.H
        entries: List[Entry] = []
        layers: List[LayerSpec] = []
        t = 0  # index into the top-level CompositeTransform
        root = "_transform._transforms"
        for i in range(self.L):
            ls = LayerSpec()
            if self.linear_transform == "lu":
                p = f"{root}.{t}._transforms"
                ls.perm_key = f"{p}.0._permutation"
                ls.lu_prefix = f"{p}.1"
                entries.append(Entry(ls.perm_key, (D,), "ibuf"))
                ntri = D * (D - 1) // 2
                entries.append(Entry(f"{p}.1.bias", (D,), "param"))
                entries.append(Entry(f"{p}.1.lower_entries", (ntri,), "param"))
                entries.append(Entry(f"{p}.1.upper_entries", (ntri,), "param"))
                entries.append(Entry(f"{p}.1.unconstrained_upper_diag", (D,), "param"))
                t += 1
            elif self.linear_transform == "permutation":
                ls.perm_key = f"{root}.{t}._permutation"
                entries.append(Entry(ls.perm_key, (D,), "ibuf"))
                t += 1
            if self.ftype == "maf":
                cp = f"{root}.{t}"
                ls.coupling_prefix = cp
                ls.identity = np.zeros(0, dtype=np.int64)
                ls.transform = np.arange(D, dtype=np.int64)
                net = f"{cp}.autoregressive_net"

                def mlin(name, n_in, n_o):
                    entries.append(Entry(f"{net}.{name}.weight", (n_o, n_in), "param"))
                    entries.append(Entry(f"{net}.{name}.bias", (n_o,), "param"))
                    entries.append(Entry(f"{net}.{name}.mask", (n_o, n_in), "fbuf"))
                    entries.append(Entry(f"{net}.{name}.degrees", (n_o,), "ibuf"))
                    ls.linears.append(LinearRef(f"{net}.{name}.weight", f"{net}.{name}.bias", n_in, n_o,
                                                f"{net}.{name}.mask"))

                mlin("initial_layer", D, H)
                for b in range(self.n_layers):
                    if self.net == "resnet":
                        mlin(f"blocks.{b}.linear_layers.0", H, H)
                        mlin(f"blocks.{b}.linear_layers.1", H, H)
                    else:
                        mlin(f"blocks.{b}.linear", H, H)
                mlin("final_layer", H, 2 * D)
                t += 1
                if self.bn_between:
                    bp = f"{root}.{t}"
                    ls.bn_prefix = bp
                    entries.append(Entry(f"{bp}.unconstrained_weight", (D,), "param"))
                    entries.append(Entry(f"{bp}.bias", (D,), "param"))
                    entries.append(Entry(f"{bp}.running_mean", (D,), "fbuf"))
                    entries.append(Entry(f"{bp}.running_var", (D,), "fbuf"))
                    t += 1
                layers.append(ls)
                continue
            cp = f"{root}.{t}"
            ls.coupling_prefix = cp
            m = self.masks[i]
            ls.identity = np.arange(D)[m <= 0].astype(np.int64)
            ls.transform = np.arange(D)[m > 0].astype(np.int64)
            d_id, d_tr = len(ls.identity), len(ls.transform)
            if d_id == 0 or d_tr == 0:
                raise ValueError("Mask must leave identity and transformed features")
            entries.append(Entry(f"{cp}.identity_features", (d_id,), "ibuf"))
            entries.append(Entry(f"{cp}.transform_features", (d_tr,), "ibuf"))
            n_out = d_tr * self.coupling_multiplier
            net = f"{cp}.transform_net"

            def lin(name, n_in, n_o):
                entries.append(Entry(f"{net}.{name}.weight", (n_o, n_in), "param"))
                entries.append(Entry(f"{net}.{name}.bias", (n_o,), "param"))
                ls.linears.append(
                    LinearRef(f"{net}.{name}.weight", f"{net}.{name}.bias", n_in, n_o)
                )

            if self.net == "mlp":
                # /root/reference/src/nessai/flows/nets.py:55-68
                lin("_input_layer", d_id, H)
                for j in range(self.n_layers - 1):
                    lin(f"_hidden_layers.{j}", H, H)
                lin("_output_layer", H, n_out)
            else:
                lin("initial_layer", d_id, H)
                for b in range(self.n_layers):
                    lin(f"blocks.{b}.linear_layers.0", H, H)
                    lin(f"blocks.{b}.linear_layers.1", H, H)
                lin("final_layer", H, n_out)
            t += 1
            if self.bn_between:
                bp = f"{root}.{t}"
                ls.bn_prefix = bp
                entries.append(Entry(f"{bp}.unconstrained_weight", (D,), "param"))
                entries.append(Entry(f"{bp}.bias", (D,), "param"))
                entries.append(Entry(f"{bp}.running_mean", (D,), "fbuf"))
                entries.append(Entry(f"{bp}.running_var", (D,), "fbuf"))
                t += 1
            layers.append(ls)
        off = 0
        for e in entries:
            if e.kind == "param":
                e.offset = off
                off += e.size
        self.n_params = off
        for e in entries:
            if e.kind == "fbuf":
                e.offset = off
                off += e.size
        self.n_theta = off
        self.entries = entries
        self.layers = layers
        self.by_key: Dict[str, Entry] = {e.key: e for e in entries}

    # ------------------------------------------------------------------- init
    def init_state(self, *, reset_bn_running_var_to: float = 0.0):
        """Fresh parameters, drawn with torch's CPU generator in the order the
        reference module tree is constructed, so ``torch.manual_seed(s)``
        followed by this call equals ``configure_model`` bit for bit.

        Returns ``(theta: np.float32[n_theta], ints: dict key -> int64 array)``.
        """
        import torch
        from torch.nn import init

        theta = np.zeros(self.n_theta, dtype=np.float32)
        ints: Dict[str, np.ndarray] = {}

        def put(key, arr):
            e = self.by_key[key]
            theta[e.offset : e.offset + e.size] = np.asarray(arr, dtype=np.float32).ravel()

        for ls in self.layers:
            if ls.perm_key is not None:
                if self.ftype == "maf" and not self.random_permutations:
                    ints[ls.perm_key] = np.arange(self.D - 1, -1, -1, dtype=np.int64)  # ReversePermutation
                else:
                    ints[ls.perm_key] = torch.randperm(self.D).numpy().astype(np.int64)
            if ls.lu_prefix is not None:
                const = np.log(np.exp(1 - self.LU_EPS) - 1)
                put(f"{ls.lu_prefix}.unconstrained_upper_diag", np.full(self.D, const))
            if self.ftype == "maf":
                self._made_masks(ls, theta, ints)
            else:
                ints[f"{ls.coupling_prefix}.identity_features"] = ls.identity.copy()
                ints[f"{ls.coupling_prefix}.transform_features"] = ls.transform.copy()
            for lr in ls.linears:
                # torch.nn.Linear.reset_parameters (stock torch)
                w = torch.empty(lr.n_out, lr.n_in)
                b = torch.empty(lr.n_out)
                init.kaiming_uniform_(w, a=math.sqrt(5))
                bound = 1 / math.sqrt(lr.n_in) if lr.n_in > 0 else 0
                init.uniform_(b, -bound, bound)
                if self.net == "resnet" and lr.weight.endswith("linear_layers.1.weight"):
                    # nflows ResidualBlock zero_initialization
                    init.uniform_(w, -1e-3, 1e-3)
                    init.uniform_(b, -1e-3, 1e-3)
                put(lr.weight, w.numpy())
                put(lr.bias, b.numpy())
            if ls.bn_prefix is not None:
                const = np.log(np.exp(1 - self.BN_EPS) - 1)
                put(f"{ls.bn_prefix}.unconstrained_weight", np.full(self.D, const))
                put(f"{ls.bn_prefix}.running_var", np.full(self.D, reset_bn_running_var_to))
        return theta, ints

    def _made_masks(self, ls: "LayerSpec", theta: np.ndarray, ints: Dict[str, np.ndarray]) -> None:
        """Degrees and masks of nflows' MADE with sequential (non-random) degrees
        (glasflow.nflows.transforms.made.MaskedLinear._get_mask_and_degrees): hidden
        units get degrees ``arange(H) % max(1, D-1) + min(1, D-1)`` and see inputs of
        degree <= their own; output (2 per feature, feature-major) ``i`` sees hidden
        units of degree < i."""
        D = self.D
        in_deg = np.arange(1, D + 1)
        for k, lr in enumerate(ls.linears):
            is_output = k == len(ls.linears) - 1
            if is_output:
                out_deg = np.repeat(np.arange(1, D + 1), lr.n_out // D)
                mask = out_deg[:, None] > in_deg[None, :]
            else:
                mx, mn = max(1, D - 1), min(1, D - 1)
                out_deg = np.arange(lr.n_out) % mx + mn
                mask = out_deg[:, None] >= in_deg[None, :]
            e = self.by_key[lr.mask]
            theta[e.offset : e.offset + e.size] = mask.astype(np.float32).ravel()
            ints[lr.mask[: -len("mask")] + "degrees"] = out_deg.astype(np.int64)
            in_deg = out_deg

    def reset_weights(self, theta: np.ndarray) -> None:
        """In-place mirror of ``model.apply(reset_weights)``
        (/root/reference/src/nessai/flows/utils.py:249-274): every ``nn.Linear``
        gets stock ``reset_parameters`` (so ResidualBlock second linears lose
        their near-zero init), flow BatchNorm gets its constants with
        ``running_var = 1``; LU and permutations are left alone."""
        import torch
        from torch.nn import init

        def put(key, arr):
            e = self.by_key[key]
            theta[e.offset : e.offset + e.size] = np.asarray(arr, dtype=np.float32).ravel()

        for ls in self.layers:
            for lr in ls.linears:
                w = torch.empty(lr.n_out, lr.n_in)
                b = torch.empty(lr.n_out)
                init.kaiming_uniform_(w, a=math.sqrt(5))
                bound = 1 / math.sqrt(lr.n_in) if lr.n_in > 0 else 0
                init.uniform_(b, -bound, bound)
                put(lr.weight, w.numpy())
                put(lr.bias, b.numpy())
            if ls.bn_prefix is not None:
                const = np.log(np.exp(1 - self.BN_EPS) - 1)
                put(f"{ls.bn_prefix}.unconstrained_weight", np.full(self.D, const))
                put(f"{ls.bn_prefix}.bias", np.zeros(self.D))
                put(f"{ls.bn_prefix}.running_mean", np.zeros(self.D))
                put(f"{ls.bn_prefix}.running_var", np.ones(self.D))

    def reset_permutations(self, theta: np.ndarray, ints: Dict[str, np.ndarray]) -> None:
        """Mirror of ``model.apply(reset_permutations)`` (flows/utils.py:277-292)."""
        import torch

        for ls in self.layers:
            if ls.perm_key is not None and not (self.ftype == "maf" and not self.random_permutations):
                ints[ls.perm_key] = torch.randperm(self.D).numpy().astype(np.int64)
            if ls.lu_prefix is not None:
                for name, val in (
                    ("bias", 0.0),
                    ("lower_entries", 0.0),
                    ("upper_entries", 0.0),
                    ("unconstrained_upper_diag", np.log(np.exp(1 - self.LU_EPS) - 1)),
                ):
                    e = self.by_key[f"{ls.lu_prefix}.{name}"]
                    theta[e.offset : e.offset + e.size] = val

    # ------------------------------------------------------------- state dict
    def get(self, theta: np.ndarray, key: str) -> np.ndarray:
        e = self.by_key[key]
        return theta[e.offset : e.offset + e.size].reshape(e.shape)

    def state_dict_numpy(self, theta, ints) -> "Dict[str, np.ndarray]":
        out = {}
        for e in self.entries:
            if e.kind == "ibuf":
                out[e.key] = np.asarray(ints[e.key], dtype=np.int64).copy()
            else:
                out[e.key] = np.array(self.get(theta, e.key), dtype=np.float32)
        return out

    def load_state_dict_numpy(self, sd, theta, ints, strict=True) -> None:
        missing = [e.key for e in self.entries if e.key not in sd]
        unexpected = [k for k in sd if k not in self.by_key]
        if strict and (missing or unexpected):
            raise RuntimeError(
                f"Error(s) in loading state_dict: missing keys {missing}, "
                f"unexpected keys {unexpected}"
            )
        for e in self.entries:
            if e.key not in sd:
                continue
            arr = np.asarray(sd[e.key])
            if tuple(arr.shape) != tuple(e.shape):
                raise RuntimeError(
                    f"size mismatch for {e.key}: {arr.shape} vs {e.shape}"
                )
            if e.kind == "ibuf":
                ints[e.key] = arr.astype(np.int64).copy()
            else:
                theta[e.offset : e.offset + e.size] = arr.astype(np.float32).ravel()

    # ---------------------------------------------------------------- folding
    def _lu_matrices(self, theta64, ls: LayerSpec):
        D = self.D
        lo = np.zeros((D, D))
        up = np.zeros((D, D))
        lo[np.tril_indices(D, k=-1)] = self.get(theta64, f"{ls.lu_prefix}.lower_entries")
        lo[np.diag_indices(D)] = 1.0
        up[np.triu_indices(D, k=1)] = self.get(theta64, f"{ls.lu_prefix}.upper_entries")
        u = self.get(theta64, f"{ls.lu_prefix}.unconstrained_upper_diag")
        diag = np.logaddexp(0.0, u) + self.LU_EPS  # softplus + eps
        up[np.diag_indices(D)] = diag
        return lo, up, diag

    def fold(self, theta: np.ndarray, ints: Dict[str, np.ndarray]) -> "FoldedFlow":
        """Fold the eval-mode flow (float64) into affine + coupling stages.

        Forward direction (x -> z) per layer: ``v = A h + b`` then coupling on
        ``[identity | transformed]`` halves; after the last coupling a final
        affine.  ``const_logdet`` collects every row-constant log|det|.
        """
        D = self.D
        th = np.asarray(theta, dtype=np.float64)
        M = np.eye(D)
        m = np.zeros(D)
        const_ld = 0.0
        stages = []
        for ls in self.layers:
            if ls.perm_key is not None:
                perm = np.asarray(ints[ls.perm_key], dtype=np.int64)
                # u_j = h_{perm[j]}
                P = np.zeros((D, D))
                P[np.arange(D), perm] = 1.0
                M, m = P @ M, P @ m
            if ls.lu_prefix is not None:
                lo, up, diag = self._lu_matrices(th, ls)
                W = lo @ up
                b = self.get(th, f"{ls.lu_prefix}.bias")
                M, m = W @ M, W @ m + b
                const_ld += float(np.sum(np.log(diag)))
            # gather: identity features first, then transformed
            order = np.concatenate([ls.identity, ls.transform])
            G = np.zeros((D, D))
            G[np.arange(D), order] = 1.0
            A, b = G @ M, G @ m
            stages.append((A, b, ls))
            # scatter back
            M, m = G.T.copy(), np.zeros(D)
            if ls.bn_prefix is not None:
                uw = self.get(th, f"{ls.bn_prefix}.unconstrained_weight")
                w = np.logaddexp(0.0, uw) + self.BN_EPS
                beta = self.get(th, f"{ls.bn_prefix}.bias")
                rm = self.get(th, f"{ls.bn_prefix}.running_mean")
                rv = self.get(th, f"{ls.bn_prefix}.running_var")
                a = w / np.sqrt(rv + self.BN_EPS)
                c = beta - a * rm
                M, m = a[:, None] * M, a * m + c
                const_ld += float(np.sum(np.log(w) - 0.5 * np.log(rv + self.BN_EPS)))
        return FoldedFlow(self, th, stages, (M, m), const_ld)


class FoldedFlow:
    """Eval-mode flow as ``affine -> (coupling -> affine)*`` in either direction."""

    def __init__(self, spec: FlowSpec, theta64, stages, final, const_logdet):
        self.spec = spec
        self.theta64 = theta64
        self.stages = stages  # [(A, b, LayerSpec)] forward order
        self.final = final  # (A_out, b_out)
        self.const_logdet = const_logdet

    # -- conditioner weights as (W (out,in), b (out,)) float64 ----------------
    def net_weights(self, ls: LayerSpec):
        sp = self.spec
        out = []
        for lr in ls.linears:
            W = sp.get(self.theta64, lr.weight)
            if lr.mask is not None:
                W = W * sp.get(self.theta64, lr.mask)
            out.append((W, sp.get(self.theta64, lr.bias)))
        return out

    def program(self, inverse: bool) -> "Program":
        """Encode one direction as the op list + float blob the kernels run."""
        sp = self.spec
        D = sp.D
        blob: List[np.ndarray] = []
        pos = 0

        def push(arr) -> int:
            nonlocal pos
            a = np.ascontiguousarray(arr, dtype=np.float64).ravel()
            # keep every block 16-byte aligned for float4 loads
            padn = (-len(a)) % 4
            if padn:
                a = np.concatenate([a, np.zeros(padn)])
            blob.append(a)
            off = pos
            pos += len(a)
            return off

        ops: List[List[int]] = []
        cur_x = BUF_X0

        def emit_linear(src, dst, Wm, bv, flags, src_off=0):
            # Wm: (N, K) ; stored k-major [K][Npad]
            N, K = Wm.shape
            Np = _pad8(N)
            Wk = np.zeros((K, Np))
            Wk[:, :N] = Wm.T
            bp = np.zeros(Np)
            bp[:N] = bv
            w_off = push(Wk)
            b_off = push(bp)
            ops.append(
                [OP_LINEAR, src, dst, src_off, K, N, Np, w_off, b_off, flags, 0, 0, 0, 0, 0, 0]
            )

        def emit_affine(A, b):
            nonlocal cur_x
            dst = BUF_X1 if cur_x == BUF_X0 else BUF_X0
            emit_linear(cur_x, dst, A, b, 0)
            cur_x = dst

        def emit_made(ls: LayerSpec, src_x, dst_x, z_x, flags):
            """One MADE pass reading its input from ``src_x``; the affine acts on ``z_x`` and
            writes ``dst_x`` (nflows MaskedAffineAutoregressiveTransform: params viewed
            (N, D, 2) = (unconstrained scale, shift); scale = softplus(u) + 1e-3)."""
            ws = self.net_weights(ls)
            W, b = ws[0]
            if sp.net == "resnet":
                emit_linear(src_x, BUF_A0, W, b, 0)
                for blk in range(sp.n_layers):
                    W0, b0 = ws[1 + 2 * blk]
                    W1, b1 = ws[2 + 2 * blk]
                    emit_linear(BUF_A0, BUF_A1, W0, b0, FLAG_IN_ACT | FLAG_OUT_ACT)
                    emit_linear(BUF_A1, BUF_A0, W1, b1, FLAG_ACCUM)
                last_src = BUF_A0
            else:
                bufs = [BUF_A0, BUF_A1]
                emit_linear(src_x, bufs[0], W, b, FLAG_OUT_ACT)
                src = bufs[0]
                for j, (Wj, bj) in enumerate(ws[1:-1]):
                    dst = bufs[(j + 1) % 2]
                    emit_linear(src, dst, Wj, bj, FLAG_OUT_ACT)
                    src = dst
                last_src = src
            Wf, bf = ws[-1]
            K = Wf.shape[1]
            # kernel layout: interleaved (shift_i, unconstrained_scale_i)
            Wi = np.zeros((2 * D, K))
            bi = np.zeros(2 * D)
            Wi[0::2], bi[0::2] = Wf[1::2], bf[1::2]
            Wi[1::2], bi[1::2] = Wf[0::2], bf[0::2]
            Np = _pad8(2 * D)
            Wk = np.zeros((K, Np))
            Wk[:, : 2 * D] = Wi.T
            bp = np.zeros(Np)
            bp[: 2 * D] = bi
            w_off = push(Wk)
            b_off = push(bp)
            ops.append([OP_COUPLING_AFFINE, last_src, dst_x, 0, K, 2 * D, Np, w_off, b_off,
                        flags | FLAG_SOFTPLUS_SCALE, z_x, 0, D, 0, 0, 0])

        def emit_coupling(ls: LayerSpec):
            nonlocal cur_x
            if sp.ftype == "maf":
                if not inverse:
                    emit_made(ls, cur_x, cur_x, cur_x, 0)
                    return
                # inverse: D sequential passes from x = 0 (AutoregressiveTransform.inverse)
                other = BUF_X1 if cur_x == BUF_X0 else BUF_X0
                emit_linear(cur_x, other, np.zeros((D, D)), np.zeros(D), 0)
                for i in range(D):
                    last = i == D - 1
                    emit_made(ls, other, other, cur_x, FLAG_INVERSE | (0 if last else FLAG_NO_LOGDET))
                cur_x = other
                return
            d_id, d_tr = len(ls.identity), len(ls.transform)
            ws = self.net_weights(ls)
            if sp.net == "mlp":
                src = cur_x
                bufs = [BUF_A0, BUF_A1]
                for j, (W, b) in enumerate(ws[:-1]):
                    dst = bufs[j % 2]
                    emit_linear(src, dst, W, b, FLAG_OUT_ACT)
                    src = dst
                last_src = src
            else:
                W, b = ws[0]
                emit_linear(cur_x, BUF_A0, W, b, 0)
                for blk in range(sp.n_layers):
                    W0, b0 = ws[1 + 2 * blk]
                    W1, b1 = ws[2 + 2 * blk]
                    emit_linear(BUF_A0, BUF_A1, W0, b0, FLAG_IN_ACT | FLAG_OUT_ACT)
                    emit_linear(BUF_A1, BUF_A0, W1, b1, FLAG_ACCUM)
                last_src = BUF_A0
            Wf, bf = ws[-1]
            K = Wf.shape[1]
            flags = FLAG_INVERSE if inverse else 0
            if sp.ftype == "realnvp":
                # outputs reordered to interleaved (shift_i, unconstrained_scale_i)
                Wi = np.zeros((2 * d_tr, K))
                bi = np.zeros(2 * d_tr)
                Wi[0::2] = Wf[:d_tr]
                bi[0::2] = bf[:d_tr]
                if sp.volume_preserving:
                    flags |= FLAG_ADDITIVE
                else:
                    Wi[1::2] = Wf[d_tr:]
                    bi[1::2] = bf[d_tr:]
                N = 2 * d_tr
                Np = _pad8(N)
                Wk = np.zeros((K, Np))
                Wk[:, :N] = Wi.T
                bp = np.zeros(Np)
                bp[:N] = bi
                op = OP_COUPLING_AFFINE
                extra = [0, 0, 0]
            else:
                # spline: per transformed feature a group of 3K-1 outputs padded
                # to G = pad8(3K-1); widths/heights pre-divided by sqrt(hidden)
                nb = sp.num_bins
                mult = 3 * nb - 1
                G = _pad8(mult)
                scale = np.ones(mult)
                scale[: 2 * nb] = 1.0 / np.sqrt(sp.H)
                N = d_tr * G
                Np = N
                Wk = np.zeros((K, Np))
                bp = np.zeros(Np)
                for f in range(d_tr):
                    Wg = Wf[f * mult : (f + 1) * mult] * scale[:, None]
                    Wk[:, f * G : f * G + mult] = Wg.T
                    bp[f * G : f * G + mult] = bf[f * mult : (f + 1) * mult] * scale
                op = OP_COUPLING_SPLINE
                extra = [nb, G, int(np.float32(sp.tail_bound).view(np.int32))]
            w_off = push(Wk)
            b_off = push(bp)
            ops.append(
                [op, last_src, cur_x, 0, K, N, Np, w_off, b_off, flags, cur_x, d_id, d_tr]
                + extra
            )

        if not inverse:
            for A, b, ls in self.stages:
                emit_affine(A, b)
                emit_coupling(ls)
            emit_affine(*self.final)
            const = self.const_logdet
        else:
            A, b = self.final
            Ai = np.linalg.inv(A)
            emit_affine(Ai, -Ai @ b)
            for A, b, ls in reversed(self.stages):
                emit_coupling(ls)
                Ai = np.linalg.inv(A)
                emit_affine(Ai, -Ai @ b)
            const = -self.const_logdet
        blob_f = np.concatenate(blob).astype(np.float32) if blob else np.zeros(0, np.float32)
        ops_i = np.asarray(ops, dtype=np.int32).reshape(-1, OP_INTS)
        return Program(
            ops=ops_i,
            blob=blob_f,
            D=D,
            H=sp.H,
            final_buf=cur_x,
            const_logdet=float(const),
            activation=sp.activation,
            inverse=inverse,
        )


@dataclass
class Program:
    ops: np.ndarray  # (n_ops, OP_INTS) int32
    blob: np.ndarray  # float32
    D: int
    H: int
    final_buf: int
    const_logdet: float
    activation: int
    inverse: bool

    @property
    def max_weight_floats(self) -> int:
        """Largest single op's staged weights (K*Npad + Npad)."""
        return int(max(o[4] * o[6] + o[6] for o in self.ops))
