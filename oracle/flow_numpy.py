"""ORACLE (test infrastructure only) -- float64 numpy restatement of the
eval-mode flows nessai assembles, evaluated layer by layer exactly as the
reference composes them (no folding), from a reference-layout ``state_dict``.

**Parity unpinned at the nflows boundary**: the arithmetic below restates
``glasflow.nflows`` (absent from /root/reference and from this image) per
SURVEY.md section 8(c); it is cross-checked against ``oracle/shims`` running
under the UNMODIFIED reference (tests/test_oracle.py) and against the
self-consistency properties the reference's own tests pin
(/root/reference/tests/test_flows/test_included_flows.py:114-154).

Assembly followed: /root/reference/src/nessai/flows/realnvp.py:175-214,
nsf.py:97-130, flows/base.py:209-287 (NFlow forward / inverse / log_prob /
sample_and_log_prob).
"""

from __future__ import annotations

import re

import numpy as np

LU_EPS = 1e-3
BN_EPS = 1e-5
MIN_BIN_WIDTH = 1e-3
MIN_BIN_HEIGHT = 1e-3
MIN_DERIVATIVE = 1e-3


def softplus(x):
    return np.logaddexp(0.0, x)


def activation(name):
    if name == "relu":
        return lambda x: np.maximum(x, 0.0)
    if name == "tanh":
        return np.tanh
    if name in ("silu", "swish"):
        # /root/reference/src/nessai/flows/utils.py:24-32
        return lambda x: x / (1.0 + np.exp(-x))
    raise ValueError(name)


def _sd64(sd):
    out = {}
    for k, v in sd.items():
        v = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v)
        out[k] = v.astype(np.float64) if v.dtype.kind == "f" else v.astype(np.int64)
    return out


class NumpyFlow:
    """``NumpyFlow(state_dict, ftype=..., net=..., activation=..., ...)``."""

    def __init__(
        self,
        state_dict,
        ftype="realnvp",
        net="resnet",
        activation_name="relu",
        volume_preserving=False,
        num_bins=8,
        tail_bound=5.0,
        hidden_features=None,
        base_var=1.0,
    ):
        # base distribution N(0, base_var I): nessai's MultivariateNormal
        # (/root/reference/src/nessai/flows/distributions.py:17-73); 1 = nflows' StandardNormal
        self.base_var = float(base_var)
        self.sd = _sd64(state_dict)
        self.ftype = ftype
        self.net = net
        self.act = activation(activation_name)
        self.additive = volume_preserving
        self.num_bins = num_bins
        self.tail_bound = tail_bound
        self.hidden_features = hidden_features
        self.transforms = self._parse()

    # ------------------------------------------------------------------ parse
    def _parse(self):
        root = "_transform._transforms"
        idx = sorted(
            {int(m.group(1)) for k in self.sd for m in [re.match(rf"{re.escape(root)}\.(\d+)\.", k)] if m}
        )
        out = []
        for t in idx:
            p = f"{root}.{t}"
            if f"{p}._transforms.0._permutation" in self.sd:
                out.append(("perm", f"{p}._transforms.0"))
                out.append(("lu", f"{p}._transforms.1"))
            elif f"{p}._permutation" in self.sd:
                out.append(("perm", p))
            elif f"{p}.identity_features" in self.sd:
                out.append(("coupling", p))
            elif f"{p}.autoregressive_net.initial_layer.weight" in self.sd:
                out.append(("maf", p))
            elif f"{p}.unconstrained_weight" in self.sd:
                out.append(("bn", p))
            else:
                raise ValueError(f"cannot classify transform {p}")
        return out

    # ------------------------------------------------------------- transforms
    def _perm(self, p, x, inverse):
        perm = self.sd[f"{p}._permutation"]
        if inverse:
            perm = np.argsort(perm)
        return x[:, perm], 0.0

    def _lu(self, p, x, inverse):
        D = x.shape[1]
        lo = np.zeros((D, D))
        up = np.zeros((D, D))
        lo[np.tril_indices(D, k=-1)] = self.sd[f"{p}.lower_entries"]
        lo[np.diag_indices(D)] = 1.0
        up[np.triu_indices(D, k=1)] = self.sd[f"{p}.upper_entries"]
        diag = softplus(self.sd[f"{p}.unconstrained_upper_diag"]) + LU_EPS
        up[np.diag_indices(D)] = diag
        b = self.sd[f"{p}.bias"]
        ld = np.sum(np.log(diag))
        if not inverse:
            return (x @ up.T) @ lo.T + b, ld
        y = np.linalg.solve(lo, (x - b).T)
        y = np.linalg.solve(up, y)
        return y.T, -ld

    def _bn(self, p, x, inverse):
        w = softplus(self.sd[f"{p}.unconstrained_weight"]) + BN_EPS
        beta = self.sd[f"{p}.bias"]
        rm, rv = self.sd[f"{p}.running_mean"], self.sd[f"{p}.running_var"]
        ld = np.sum(np.log(w) - 0.5 * np.log(rv + BN_EPS))
        if not inverse:
            return w * ((x - rm) / np.sqrt(rv + BN_EPS)) + beta, ld
        return np.sqrt(rv + BN_EPS) * ((x - beta) / w) + rm, -ld

    def _net(self, p, h):
        sd, act = self.sd, self.act
        n = f"{p}.transform_net"

        def lin(name, v):
            return v @ sd[f"{n}.{name}.weight"].T + sd[f"{n}.{name}.bias"]

        if self.net == "mlp":
            # /root/reference/src/nessai/flows/nets.py:113-126
            h = act(lin("_input_layer", h))
            j = 0
            while f"{n}._hidden_layers.{j}.weight" in sd:
                h = act(lin(f"_hidden_layers.{j}", h))
                j += 1
            return lin("_output_layer", h)
        h = lin("initial_layer", h)
        b = 0
        while f"{n}.blocks.{b}.linear_layers.0.weight" in sd:
            t = act(h)
            t = lin(f"blocks.{b}.linear_layers.0", t)
            t = act(t)
            t = lin(f"blocks.{b}.linear_layers.1", t)
            h = h + t
            b += 1
        return lin("final_layer", h)

    def _coupling(self, p, x, inverse):
        idf = self.sd[f"{p}.identity_features"]
        trf = self.sd[f"{p}.transform_features"]
        ident, tr = x[:, idf], x[:, trf]
        params = self._net(p, ident)
        d_tr = len(trf)
        if self.ftype == "realnvp":
            if self.additive:
                shift, scale = params, np.ones_like(params)
            else:
                shift = params[:, :d_tr]
                scale = 1.0 / (1.0 + np.exp(-(params[:, d_tr:] + 2.0))) + 1e-3
            if not inverse:
                tr2, ld = tr * scale + shift, np.sum(np.log(scale), axis=1)
            else:
                tr2, ld = (tr - shift) / scale, -np.sum(np.log(scale), axis=1)
        else:
            tr2, ld = self._spline(tr, params.reshape(len(x), d_tr, -1), inverse)
            ld = np.sum(ld, axis=1)
        out = np.empty_like(x)
        out[:, idf] = ident
        out[:, trf] = tr2
        return out, ld

    # -------------------------------------------------------------- RQ spline
    def _spline(self, x, p, inverse):
        K, B = self.num_bins, self.tail_bound
        uw, uh, ud = p[..., :K], p[..., K : 2 * K], p[..., 2 * K :]
        if self.hidden_features is not None:
            uw = uw / np.sqrt(self.hidden_features)
            uh = uh / np.sqrt(self.hidden_features)
        inside = (x >= -B) & (x <= B)
        const = np.log(np.exp(1 - MIN_DERIVATIVE) - 1)
        ud = np.concatenate(
            [np.full(ud.shape[:-1] + (1,), const), ud, np.full(ud.shape[:-1] + (1,), const)], -1
        )

        def knots(u, min_size):
            e = np.exp(u - u.max(-1, keepdims=True))
            sm = e / e.sum(-1, keepdims=True)
            w = min_size + (1 - min_size * K) * sm
            cw = np.concatenate([np.zeros(w.shape[:-1] + (1,)), np.cumsum(w, -1)], -1)
            cw = 2 * B * cw - B
            cw[..., 0] = -B
            cw[..., -1] = B
            return cw, cw[..., 1:] - cw[..., :-1]

        cw, w = knots(uw, MIN_BIN_WIDTH)
        ch, h = knots(uh, MIN_BIN_HEIGHT)
        d = MIN_DERIVATIVE + softplus(ud)
        loc = ch if inverse else cw
        loc = loc.copy()
        loc[..., -1] += 1e-6
        xc = np.clip(x, -B, B)
        b = np.sum(xc[..., None] >= loc, -1) - 1
        b = np.clip(b, 0, K - 1)

        def g(a):
            return np.take_along_axis(a, b[..., None], -1)[..., 0]

        icw, iw, ich, ih = g(cw), g(w), g(ch), g(h)
        delta = h / w
        idl, id0, id1 = g(delta), g(d[..., :-1]), g(d[..., 1:])
        if inverse:
            a = (xc - ich) * (id0 + id1 - 2 * idl) + ih * (idl - id0)
            bb = ih * id0 - (xc - ich) * (id0 + id1 - 2 * idl)
            c = -idl * (xc - ich)
            disc = bb**2 - 4 * a * c
            # the reference asserts disc >= 0 for the whole batch; here rows
            # that violate it come out NaN (row-level semantics)
            root = (2 * c) / (-bb - np.sqrt(disc))
            out = root * iw + icw
            t1m = root * (1 - root)
            den = idl + (id0 + id1 - 2 * idl) * t1m
            dnum = idl**2 * (id1 * root**2 + 2 * idl * t1m + id0 * (1 - root) ** 2)
            ld = -(np.log(dnum) - 2 * np.log(den))
        else:
            th = (xc - icw) / iw
            t1m = th * (1 - th)
            num = ih * (idl * th**2 + id0 * t1m)
            den = idl + (id0 + id1 - 2 * idl) * t1m
            out = ich + num / den
            dnum = idl**2 * (id1 * th**2 + 2 * idl * t1m + id0 * (1 - th) ** 2)
            ld = np.log(dnum) - 2 * np.log(den)
        out = np.where(inside, out, x)
        ld = np.where(inside, ld, 0.0)
        return out, ld

    # ------------------------------------------- masked affine autoregressive
    def _made(self, p, h):
        """nflows MADE (/root/reference/src/nessai/flows/maf.py:84-98 ->
        MaskedAffineAutoregressiveTransform): masked linears, residual blocks
        ``h += lin1(act(lin0(act(h))))`` or feed-forward blocks ``act(lin(h))``."""
        sd, act = self.sd, self.act
        n = f"{p}.autoregressive_net"

        def lin(name, v):
            return v @ (sd[f"{n}.{name}.weight"] * sd[f"{n}.{name}.mask"]).T + sd[f"{n}.{name}.bias"]

        h = lin("initial_layer", h)
        if f"{n}.blocks.0.linear_layers.0.weight" in sd:
            b = 0
            while f"{n}.blocks.{b}.linear_layers.0.weight" in sd:
                t = lin(f"blocks.{b}.linear_layers.0", act(h))
                h = h + lin(f"blocks.{b}.linear_layers.1", act(t))
                b += 1
        else:
            h = act(h)
            b = 0
            while f"{n}.blocks.{b}.linear.weight" in sd:
                h = act(lin(f"blocks.{b}.linear", h))
                b += 1
        return lin("final_layer", h)

    def _maf(self, p, x, inverse):
        """params viewed (N, D, 2) = (unconstrained scale, shift); scale = softplus(u) + 1e-3;
        inverse = D sequential passes from zeros (AutoregressiveTransform.inverse)."""
        D = x.shape[1]

        def scale_shift(inp):
            prm = self._made(p, inp).reshape(len(inp), D, 2)
            return softplus(prm[..., 0]) + 1e-3, prm[..., 1]

        if not inverse:
            s, t = scale_shift(x)
            return s * x + t, np.sum(np.log(s), axis=1)
        out = np.zeros_like(x)
        for _ in range(D):
            s, t = scale_shift(out)
            out = (x - t) / s
        return out, -np.sum(np.log(s), axis=1)

    # ---------------------------------------------------------------- public
    def _apply(self, kind, p, x, inverse):
        return {
            "perm": self._perm,
            "lu": self._lu,
            "bn": self._bn,
            "coupling": self._coupling,
            "maf": self._maf,
        }[kind](p, x, inverse)

    def forward(self, x):
        """x -> (z, log|det J|)  (flows/base.py:209-214)."""
        x = np.asarray(x, dtype=np.float64)
        ld = np.zeros(len(x))
        for kind, p in self.transforms:
            x, l = self._apply(kind, p, x, False)
            ld = ld + l
        return x, ld

    def inverse(self, z):
        """z -> (x, log|det J_inverse|)  (flows/base.py:216-221)."""
        z = np.asarray(z, dtype=np.float64)
        ld = np.zeros(len(z))
        for kind, p in reversed(self.transforms):
            z, l = self._apply(kind, p, z, True)
            ld = ld + l
        return z, ld

    def base_log_prob(self, z):
        D = z.shape[1]
        var = getattr(self, "base_var", 1.0)  # (also callable as NumpyFlow.base_log_prob(None, z))
        return -(0.5 / var) * np.sum(z**2, axis=1) - 0.5 * D * np.log(2 * np.pi * var)

    def forward_and_log_prob(self, x):
        z, ld = self.forward(x)
        return z, self.base_log_prob(z) + ld

    def log_prob(self, x):
        return self.forward_and_log_prob(x)[1]

    def sample_and_log_prob(self, z):
        """``FlowModel.sample_and_log_prob(z=z)`` semantics
        (/root/reference/src/nessai/flowmodel/base.py:939-944)."""
        x, ld = self.inverse(z)
        return x, self.base_log_prob(np.asarray(z, dtype=np.float64)) - ld
