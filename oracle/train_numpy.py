"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

float64 numpy restatement of ONE optimisation step of ``FlowModel._train``
(/root/reference/src/nessai/flowmodel/base.py:365-452) for a RealNVP flow:

* train-mode forward ``loss = -mean(log_prob(x))`` (or the weighted loss,
  ``base.py:404-407``) through permutation -> LU (uncached) -> affine/additive
  coupling (MLP or ResidualNet conditioner) -> BatchNorm with BATCH statistics
  (unbiased variance) -- the nflows arithmetic restated in SURVEY.md 8(c);
* the hand-derived backward pass (this is what the CUDA training kernels
  implement, phase for phase);
* ``clip_grad_norm_`` + Adam / AdamW exactly as stock torch defines them
  (SURVEY.md 8a, row a16).

It is pinned by ``tests/test_train_oracle.py`` against torch autograd through
the reference's own module tree on the glasflow shim.
"""

from __future__ import annotations

import math

import numpy as np

ACT_RELU, ACT_TANH, ACT_SILU = 0, 1, 2
LU_EPS, BN_EPS, BN_MOMENTUM = 1e-3, 1e-5, 0.1


def _softplus(x):
    return np.logaddexp(0.0, x)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _act(kind, z):
    if kind == ACT_RELU:
        return np.maximum(z, 0.0)
    if kind == ACT_TANH:
        return np.tanh(z)
    return z * _sigmoid(z)


def _dact(kind, z):
    if kind == ACT_RELU:
        return (z > 0).astype(z.dtype)
    if kind == ACT_TANH:
        return 1.0 - np.tanh(z) ** 2
    s = _sigmoid(z)
    return s * (1.0 + z * (1.0 - s))


def net_ops(spec, ls):
    """The conditioner as a list of linear ops over per-row buffers.

    Buffer 0 is the identity half; op j computes
    ``buf[out] = W_j f(buf[in]) + b_j (+ buf[res])`` with ``f`` = activation if
    ``pre_act`` else identity.  MLP (nessai/flows/nets.py:83-126): activation on
    every hidden output.  ResidualNet (nflows): no activation after the initial
    layer nor before the final layer; each block ``h += lin1(act(lin0(act(h))))``.
    """
    ops = []
    n = len(ls.linears)
    if spec.net == "mlp":
        for j in range(n):
            ops.append(dict(lin=ls.linears[j], inp=j, out=j + 1, res=-1, pre_act=j > 0))
        return ops, n + 1
    ops.append(dict(lin=ls.linears[0], inp=0, out=1, res=-1, pre_act=False))
    h = 1
    nb = 2
    for b in range(spec.n_layers):
        ops.append(dict(lin=ls.linears[1 + 2 * b], inp=h, out=nb, res=-1, pre_act=True))
        ops.append(dict(lin=ls.linears[2 + 2 * b], inp=nb, out=nb + 1, res=h, pre_act=True))
        h = nb + 1
        nb += 2
    ops.append(dict(lin=ls.linears[-1], inp=h, out=nb, res=-1, pre_act=False))
    return ops, nb + 1


MIN_BIN_WIDTH = MIN_BIN_HEIGHT = MIN_DERIVATIVE = 1e-3


def _spline_knots(u, B):
    """softmax -> min size -> cumulative knots with both ends forced to -B / +B
    (nflows rational_quadratic_spline).  Returns (p, knots[..., K+1])."""
    K = u.shape[-1]
    e = np.exp(u - u.max(-1, keepdims=True))
    p = e / e.sum(-1, keepdims=True)
    w = MIN_BIN_WIDTH + (1 - MIN_BIN_WIDTH * K) * p
    s = np.concatenate([np.zeros(w.shape[:-1] + (1,)), np.cumsum(w, -1)], -1) * 2 * B - B
    s[..., 0] = -B
    s[..., -1] = B
    return p, s


def spline_forward(x, params, B, hidden):
    """Rational-quadratic spline coupling, forward direction, linear tails.
    x (N, F); params (N, F, 3K-1).  Returns (y, logdet (N, F), cache)."""
    K = (params.shape[-1] + 1) // 3
    uw = params[..., :K] / np.sqrt(hidden)
    uh = params[..., K : 2 * K] / np.sqrt(hidden)
    ud = params[..., 2 * K :]
    pw, sw = _spline_knots(uw, B)
    ph, sh = _spline_knots(uh, B)
    const = np.log(np.exp(1 - MIN_DERIVATIVE) - 1)
    udp = np.concatenate([np.full(ud.shape[:-1] + (1,), const), ud, np.full(ud.shape[:-1] + (1,), const)], -1)
    d = MIN_DERIVATIVE + _softplus(udp)
    inside = (x >= -B) & (x <= B)
    loc = sw.copy()
    loc[..., -1] += 1e-6
    k = np.clip(np.sum(np.clip(x, -B, B)[..., None] >= loc, -1) - 1, 0, K - 1)

    def g(a, idx):
        return np.take_along_axis(a, idx[..., None], -1)[..., 0]

    a, W = g(sw, k), g(sw, k + 1) - g(sw, k)
    c, Hh = g(sh, k), g(sh, k + 1) - g(sh, k)
    d0, d1 = g(d, k), g(d, k + 1)
    th = (x - a) / W
    dl = Hh / W
    t = th * (1 - th)
    Nn = Hh * (dl * th**2 + d0 * t)
    Dn = dl + (d0 + d1 - 2 * dl) * t
    Q = d1 * th**2 + 2 * dl * t + d0 * (1 - th) ** 2
    y = np.where(inside, c + Nn / Dn, x)
    ld = np.where(inside, 2 * np.log(dl) + np.log(Q) - 2 * np.log(Dn), 0.0)
    cache = dict(K=K, B=B, hidden=hidden, pw=pw, ph=ph, udp=udp, k=k, inside=inside, W=W, Hh=Hh, d0=d0, d1=d1,
                 th=th, dl=dl, t=t, Nn=Nn, Dn=Dn, Q=Q)
    return y, ld, cache


def spline_backward(gy, gl, cache):
    """Gradients of sum(gy * y + gl * logdet) w.r.t. x and the raw conditioner outputs."""
    C = cache
    K, B = C["K"], C["B"]
    W, Hh, d0, d1, th, dl, t, Nn, Dn, Q = (C[n] for n in ("W", "Hh", "d0", "d1", "th", "dl", "t", "Nn", "Dn", "Q"))
    inside = C["inside"]
    gN = gy / Dn
    gD = -gy * Nn / Dn**2
    dDn_dth = (d0 + d1 - 2 * dl) * (1 - 2 * th)
    dQ_dth = 2 * d1 * th + 2 * dl * (1 - 2 * th) - 2 * d0 * (1 - th)
    gth = gN * Hh * (2 * dl * th + d0 * (1 - 2 * th)) + gD * dDn_dth + gl * (dQ_dth / Q - 2 * dDn_dth / Dn)
    gdl = gN * Hh * th**2 + gD * (1 - 2 * t) + gl * (2 / dl + 2 * t / Q - 2 * (1 - 2 * t) / Dn)
    gH = gN * (dl * th**2 + d0 * t) + gdl / W
    gd0 = gN * Hh * t + gD * t + gl * ((1 - th) ** 2 / Q - 2 * t / Dn)
    gd1 = gD * t + gl * (th**2 / Q - 2 * t / Dn)
    gx_in = gth / W
    ga = -gth / W
    gW = -gth * th / W - gdl * dl / W
    gc = gy
    z = lambda v: np.where(inside, v, 0.0)  # noqa: E731
    ga, gW, gc, gH, gd0, gd1 = z(ga), z(gW), z(gc), z(gH), z(gd0), z(gd1)
    gx = np.where(inside, gx_in, gy)
    k = C["k"]
    shape = k.shape + (K + 1,)

    def knot_grads(g_left, g_size):
        """d/d knots s_j from the left-knot position and the bin size s_{k+1} - s_k; the two
        end knots are constants."""
        gs = np.zeros(shape)
        np.put_along_axis(gs, k[..., None], np.take_along_axis(gs, k[..., None], -1) + (g_left - g_size)[..., None], -1)
        np.put_along_axis(gs, (k + 1)[..., None], np.take_along_axis(gs, (k + 1)[..., None], -1) + g_size[..., None], -1)
        gs[..., 0] = 0.0
        gs[..., -1] = 0.0
        # s_j = -B + 2B sum_{i<j} w_i  ->  d/dw_i = 2B sum_{j>i} gs_j (j <= K-1)
        rev = np.cumsum(gs[..., ::-1], -1)[..., ::-1]
        return 2 * B * rev[..., 1:]

    def softmax_back(p, gwt, minsize):
        gp = (1 - minsize * K) * gwt
        return p * (gp - np.sum(p * gp, -1, keepdims=True))

    guw = softmax_back(C["pw"], knot_grads(ga, gW), MIN_BIN_WIDTH) / np.sqrt(C["hidden"])
    guh = softmax_back(C["ph"], knot_grads(gc, gH), MIN_BIN_HEIGHT) / np.sqrt(C["hidden"])
    gd = np.zeros(shape)
    np.put_along_axis(gd, k[..., None], gd0[..., None], -1)
    np.put_along_axis(gd, (k + 1)[..., None], np.take_along_axis(gd, (k + 1)[..., None], -1) + gd1[..., None], -1)
    gud = (gd * _sigmoid(C["udp"]))[..., 1:-1]
    return gx, np.concatenate([guw, guh, gud], -1)


class TrainStepOracle:
    """Forward + backward + optimiser step on a flat float64 ``theta``."""

    def __init__(self, spec, ints):
        self.spec = spec
        self.ints = ints

    # ------------------------------------------------------------------ utils
    def _get(self, theta, key):
        e = self.spec.by_key[key]
        return theta[e.offset : e.offset + e.size].reshape(e.shape)

    def _lu(self, theta, ls):
        D = self.spec.D
        lo = np.eye(D)
        up = np.zeros((D, D))
        lo[np.tril_indices(D, -1)] = self._get(theta, f"{ls.lu_prefix}.lower_entries")
        up[np.triu_indices(D, 1)] = self._get(theta, f"{ls.lu_prefix}.upper_entries")
        ud = self._get(theta, f"{ls.lu_prefix}.unconstrained_upper_diag")
        diag = _softplus(ud) + LU_EPS
        up[np.diag_indices(D)] = diag
        return lo, up, diag, ud

    # ---------------------------------------------------------- loss and grad
    def loss_and_grad(self, theta, x, weights=None, update_running=True):
        """Returns ``(loss, grad[n_params], new_theta_buffers_applied_in_place)``.

        ``theta`` (float64, n_theta) -- the BatchNorm running statistics inside
        it are EMA-updated in place when ``update_running``.
        """
        sp = self.spec
        D = sp.D
        x = np.asarray(x, dtype=np.float64)
        B = x.shape[0]
        c = np.full(B, 1.0 / B) if weights is None else np.asarray(weights, np.float64) / np.sum(weights)
        act = sp.activation
        saved = []
        h = x
        ld_rows = np.zeros(B)
        ld_const = 0.0
        for ls in sp.layers:
            S = {}
            if ls.perm_key is not None:
                perm = np.asarray(self.ints[ls.perm_key])
                h = h[:, perm]
                S["perm"] = perm
            S["h1"] = h
            if ls.lu_prefix is not None:
                lo, up, diag, ud = self._lu(theta, ls)
                W = lo @ up
                h = h @ W.T + self._get(theta, f"{ls.lu_prefix}.bias")
                ld_const += np.sum(np.log(diag))
                S.update(lo=lo, up=up, diag=diag, ud=ud, W=W)
            S["h2"] = h
            maf = sp.ftype == "maf"
            ident, tr = (h, h) if maf else (h[:, ls.identity], h[:, ls.transform])
            ops, nbuf = net_ops(sp, ls)
            bufs = [None] * nbuf
            bufs[0] = ident
            def weight(lr):
                W = self._get(theta, lr.weight)
                return W * self._get(theta, lr.mask) if getattr(lr, "mask", None) else W

            for op in ops:
                a = bufs[op["inp"]]
                if op["pre_act"]:
                    a = _act(act, a)
                o = a @ weight(op["lin"]).T + self._get(theta, op["lin"].bias)
                if op["res"] >= 0:
                    o = o + bufs[op["res"]]
                bufs[op["out"]] = o
            p = bufs[-1]
            d_tr = tr.shape[1]
            if maf:
                # MaskedAffineAutoregressiveTransform: params (B, D, 2) = (unconstrained scale, shift)
                s = _softplus(p[:, 0::2]) + 1e-3
                tr2 = tr * s + p[:, 1::2]
                ld_rows = ld_rows + np.sum(np.log(s), axis=1)
            elif sp.ftype == "nsf":
                s = None
                tr2, ldf, S["spline"] = spline_forward(tr, p.reshape(B, d_tr, -1), sp.tail_bound, sp.H)
                ld_rows = ld_rows + ldf.sum(1)
            elif sp.volume_preserving:
                s = np.ones_like(tr)
                tr2 = tr + p
            else:
                s = _sigmoid(p[:, d_tr:] + 2.0) + 1e-3
                tr2 = tr * s + p[:, :d_tr]
                ld_rows = ld_rows + np.sum(np.log(s), axis=1)
            y = np.empty_like(h)
            if not maf:
                y[:, ls.identity] = ident
            y[:, ls.transform] = tr2
            S.update(ops=ops, bufs=bufs, s=s, tr=tr, y=y)
            h = y
            if ls.bn_prefix is not None:
                uw = self._get(theta, f"{ls.bn_prefix}.unconstrained_weight")
                w = _softplus(uw) + BN_EPS
                beta = self._get(theta, f"{ls.bn_prefix}.bias")
                mean = y.mean(0)
                var = y.var(0, ddof=1)
                sig = np.sqrt(var + BN_EPS)
                xh = (y - mean) / sig
                h = w * xh + beta
                ld_const += np.sum(np.log(w) - 0.5 * np.log(var + BN_EPS))
                S.update(uw=uw, w=w, sig=sig, xh=xh, mean=mean, var=var)
                if update_running:
                    rm = self._get(theta, f"{ls.bn_prefix}.running_mean")
                    rv = self._get(theta, f"{ls.bn_prefix}.running_var")
                    rm *= 1 - BN_MOMENTUM
                    rm += BN_MOMENTUM * mean
                    rv *= 1 - BN_MOMENTUM
                    rv += BN_MOMENTUM * var
            saved.append(S)
        z = h
        bvar = float(getattr(sp, "base_var", 1.0))  # N(0, var I) base (flows/distributions.py:17-73)
        logp = -(0.5 / bvar) * np.sum(z * z, axis=1) - 0.5 * D * math.log(2 * math.pi * bvar) + ld_rows + ld_const
        loss = -np.sum(c * logp)

        # ------------------------------------------------------------ backward
        g = np.zeros(sp.n_params)

        def gput(key, val):
            e = sp.by_key[key]
            g[e.offset : e.offset + e.size] += np.asarray(val).ravel()

        dh = c[:, None] * z / bvar  # d loss / d z
        g_ld = -1.0  # sum over rows of d loss / d ld_row (= -sum c)
        for ls, S in zip(reversed(sp.layers), reversed(saved)):
            if ls.bn_prefix is not None:
                w, sig, xh = S["w"], S["sig"], S["xh"]
                S1 = dh.sum(0)
                S2 = (dh * xh).sum(0)
                gput(f"{ls.bn_prefix}.bias", S1)
                dw = S2 + g_ld / w
                gput(f"{ls.bn_prefix}.unconstrained_weight", dw * _sigmoid(S["uw"]))
                dxh = dh * w
                dy = (dxh - xh * (w * S2 + g_ld) / (B - 1) - (w * S1) / B) / sig
            else:
                dy = dh
            # coupling
            maf = sp.ftype == "maf"
            dtr2 = dy[:, ls.transform]
            did = np.zeros_like(dy) if maf else dy[:, ls.identity].copy()
            s, tr = S["s"], S["tr"]
            d_tr = tr.shape[1]
            if maf:
                ds = dtr2 * tr + (-c)[:, None] / s
                dp = np.empty((B, 2 * d_tr))
                dp[:, 0::2] = ds * _sigmoid(S["bufs"][-1][:, 0::2])
                dp[:, 1::2] = dtr2
                dtr = dtr2 * s
            elif sp.ftype == "nsf":
                dtr, dpp = spline_backward(dtr2, np.broadcast_to((-c)[:, None], dtr2.shape), S["spline"])
                dp = dpp.reshape(B, -1)
            elif sp.volume_preserving:
                dp = dtr2
                dtr = dtr2
            else:
                dshift = dtr2
                ds = dtr2 * tr + (-c)[:, None] / s
                sg = s - 1e-3
                du = ds * sg * (1.0 - sg)
                dp = np.concatenate([dshift, du], axis=1)
                dtr = dtr2 * s
            ops, bufs = S["ops"], S["bufs"]
            gb = [np.zeros_like(b) for b in bufs]
            gb[-1] = dp
            for op in reversed(ops):
                delta = gb[op["out"]]
                a_pre = bufs[op["inp"]]
                a = _act(act, a_pre) if op["pre_act"] else a_pre
                gW = delta.T @ a
                if getattr(op["lin"], "mask", None):
                    gW = gW * self._get(theta, op["lin"].mask)
                gput(op["lin"].weight, gW)
                gput(op["lin"].bias, delta.sum(0))
                da = delta @ weight(op["lin"])
                if op["pre_act"]:
                    da = da * _dact(act, a_pre)
                gb[op["inp"]] = gb[op["inp"]] + da
                if op["res"] >= 0:
                    gb[op["res"]] = gb[op["res"]] + delta
            did += gb[0]
            if maf:
                dh2 = dtr + did
            else:
                dh2 = np.empty_like(dy)
                dh2[:, ls.identity] = did
                dh2[:, ls.transform] = dtr
            if ls.lu_prefix is not None:
                lo, up, diag, ud, W = S["lo"], S["up"], S["diag"], S["ud"], S["W"]
                gput(f"{ls.lu_prefix}.bias", dh2.sum(0))
                dW = dh2.T @ S["h1"]
                dlo = dW @ up.T
                dup = lo.T @ dW
                gput(f"{ls.lu_prefix}.lower_entries", dlo[np.tril_indices(D, -1)])
                gput(f"{ls.lu_prefix}.upper_entries", dup[np.triu_indices(D, 1)])
                ddiag = np.diag(dup) + g_ld / diag
                gput(f"{ls.lu_prefix}.unconstrained_upper_diag", ddiag * _sigmoid(ud))
                dh1 = dh2 @ W
            else:
                dh1 = dh2
            if ls.perm_key is not None:
                dh = np.empty_like(dh1)
                dh[:, S["perm"]] = dh1
            else:
                dh = dh1
        return loss, g

    # -------------------------------------------------------------- optimiser
    @staticmethod
    def clip_(g, max_norm):
        """torch.nn.utils.clip_grad_norm_ over all parameters."""
        total = math.sqrt(float(np.sum(g * g)))
        coef = min(1.0, max_norm / (total + 1e-6))
        g *= coef
        return total

    @staticmethod
    def adam_step_(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decoupled=True):
        """torch.optim.AdamW (decoupled) / torch.optim.Adam (L2) single-tensor update;
        ``t`` is the 1-based step count."""
        if decoupled:
            p *= 1.0 - lr * weight_decay
        elif weight_decay:
            g = g + weight_decay * p
        m *= beta1
        m += (1 - beta1) * g
        v *= beta2
        v += (1 - beta2) * g * g
        bc1 = 1 - beta1 ** t
        bc2 = 1 - beta2 ** t
        denom = np.sqrt(v) / math.sqrt(bc2) + eps
        p -= (lr / bc1) * m / denom
