"""Packed description of a flow for the fused training kernels (``csrc/train.cuh``).

The reference trains through torch autograd over its Module tree
(/root/reference/src/nessai/flowmodel/base.py:365-452).  Here the architecture and
the flat parameter layout of a :class:`~nessai_b200.spec.FlowSpec` are packed into
the int32 ``TrPlan`` struct the kernels read; the field order below must match
``struct TrLinear / TrLayer / TrPlan`` exactly.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np

from .spec import FlowSpec

TR_R = 16
TR_MAXL = 16
TR_MAXBUF = 12
TR_MAXLIN = 12
TR_MAXD = 64
TR_LAYER_INTS = 20 + 2 * TR_MAXBUF + 8 * TR_MAXLIN
TR_PLAN_INTS = 20 + TR_MAXL * TR_LAYER_INTS


class TrainPlanUnsupported(NotImplementedError):
    pass


def conditioner_ops(spec: FlowSpec, ls):
    """Linear ops of the conditioner over per-row buffers (buffer 0 = identity half):
    ``buf[out] = W f(buf[in]) + b (+ buf[res])``, ``f`` = activation when ``pre_act``.
    MLP: /root/reference/src/nessai/flows/nets.py:83-126; ResidualNet: nflows."""
    ops = []
    n = len(ls.linears)
    if spec.net == "mlp":
        for j in range(n):
            ops.append((ls.linears[j], j, j + 1, -1, int(j > 0)))
        return ops, n + 1
    ops.append((ls.linears[0], 0, 1, -1, 0))
    h, nb = 1, 2
    for b in range(spec.n_layers):
        ops.append((ls.linears[1 + 2 * b], h, nb, -1, 1))
        ops.append((ls.linears[2 + 2 * b], nb, nb + 1, h, 1))
        h = nb + 1
        nb += 2
    ops.append((ls.linears[-1], h, nb, -1, 0))
    return ops, nb + 1


def param_mask(spec: FlowSpec):
    """``(n_params,)`` float32 multiplier for MADE flows: the position of every masked-linear
    weight, to be filled with its mask (see :func:`fill_param_mask`); None for other flows."""
    if spec.ftype != "maf":
        return None
    return [(spec.by_key[lr.weight].offset, spec.by_key[lr.mask].offset - spec.n_params, lr.n_in * lr.n_out)
            for ls in spec.layers for lr in ls.linears]


def build_train_plan(spec: FlowSpec, ints: dict) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns ``(plan int32[TR_PLAN_INTS], itab int32[...], reduce_idx int32[...])``."""
    if spec.ftype == "nsf" and not (2 <= spec.num_bins <= 16):
        raise TrainPlanUnsupported("fused training kernels cover spline flows with 2..16 bins")
    D, L = spec.D, spec.L
    if D > TR_MAXD or L > TR_MAXL:
        raise TrainPlanUnsupported(f"fused training kernels cover D <= {TR_MAXD}, n_blocks <= {TR_MAXL}")
    off = lambda key: spec.by_key[key].offset  # noqa: E731
    boff = lambda key: spec.by_key[key].offset - spec.n_params  # noqa: E731
    itab = []
    reduce_idx = []
    layers = np.zeros((TR_MAXL, TR_LAYER_INTS), dtype=np.int64)
    ws_off = 0
    n_part = spec.n_params
    max_dim, vals_floats, wmax, max_in = D, 0, 0, D
    for l, ls in enumerate(spec.layers):
        row = layers[l]
        row[:] = 0
        perm_off = -1
        if ls.perm_key is not None:
            perm_off = len(itab)
            itab.extend(int(v) for v in np.asarray(ints[ls.perm_key]))
        lu = [-1, -1, -1, -1]
        lu_part = -1
        if ls.lu_prefix is not None:
            lu = [off(f"{ls.lu_prefix}.{k}") for k in ("bias", "lower_entries", "upper_entries", "unconstrained_upper_diag")]
            lu_part = n_part
            n_part += D * D
            reduce_idx.extend(range(lu[0], lu[0] + D))
        bn = [-1, -1, -1, -1]
        if ls.bn_prefix is not None:
            bn = [off(f"{ls.bn_prefix}.unconstrained_weight"), off(f"{ls.bn_prefix}.bias"),
                  boff(f"{ls.bn_prefix}.running_mean"), boff(f"{ls.bn_prefix}.running_var")]
        # MAF: every feature is conditioner input AND transformed
        identity = ls.transform if spec.ftype == "maf" else ls.identity
        id_off = len(itab)
        itab.extend(int(v) for v in identity)
        tr_off = len(itab)
        itab.extend(int(v) for v in ls.transform)
        d_id, d_tr = len(identity), len(ls.transform)
        ops, n_buf = conditioner_ops(spec, ls)
        if len(ops) > TR_MAXLIN or n_buf > TR_MAXBUF:
            raise TrainPlanUnsupported("conditioner too deep for the fused training kernels")
        buf_dim = [0] * n_buf
        buf_dim[0] = d_id
        for lr, i_in, i_out, i_res, pre in ops:
            if buf_dim[i_in] != lr.n_in:
                raise AssertionError("conditioner wiring")
            buf_dim[i_out] = lr.n_out
        buf_off = list(np.concatenate([[0], np.cumsum(buf_dim)[:-1]]))
        rec = D + sum(buf_dim[1:]) + D
        row[0] = perm_off
        row[1:5] = lu
        row[5:9] = bn
        row[9:13] = [id_off, tr_off, d_id, d_tr]
        row[13:16] = [len(ops), n_buf, rec]
        row[16] = ws_off
        row[17] = lu_part
        row[20 : 20 + n_buf] = buf_dim
        row[20 + TR_MAXBUF : 20 + TR_MAXBUF + n_buf] = buf_off
        base = 20 + 2 * TR_MAXBUF
        for j, (lr, i_in, i_out, i_res, pre) in enumerate(ops):
            w, b = off(lr.weight), off(lr.bias)
            row[base + 8 * j : base + 8 * j + 8] = [w, b, lr.n_in, lr.n_out, i_in, i_out, i_res, pre]
            reduce_idx.extend(range(w, w + lr.n_in * lr.n_out))
            reduce_idx.extend(range(b, b + lr.n_out))
            wmax = max(wmax, lr.n_in * (min(lr.n_out, 64) | 1))  # one unit = <= 64 output columns
            max_in = max(max_in, lr.n_in)
        ws_off += rec
        max_dim = max(max_dim, max(buf_dim))
        vals_floats = max(vals_floats, sum(buf_dim))
    head = np.zeros(20, dtype=np.int64)
    head[16] = int(np.float32(getattr(spec, "base_var", 1.0)).view(np.int32))
    head[:16] = [
        D, L, spec.activation, 2 if spec.ftype == "maf" else int(spec.volume_preserving),
        spec.n_params, n_part, ws_off, max_dim,
        vals_floats, wmax, len(itab), len(reduce_idx),
        spec.num_bins if spec.ftype == "nsf" else 0,
        int(np.float32(spec.tail_bound).view(np.int32)) if spec.ftype == "nsf" else 0,
        spec.H, max_in,
    ]
    plan = np.concatenate([head, layers.ravel()]).astype(np.int32)
    assert plan.size == TR_PLAN_INTS
    return (
        plan,
        np.asarray(itab if itab else [0], dtype=np.int32)[: max(len(itab), 1)],
        np.asarray(reduce_idx, dtype=np.int32),
    )
