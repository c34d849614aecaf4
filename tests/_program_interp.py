"""Test helper: numpy interpreter of the folded op program (host-logic check of
``nessai_b200.spec`` -- mirrors what csrc/flow_generic.cu executes)."""

import numpy as np

from nessai_b200 import spec as S


def _act(kind, x):
    if kind == S.ACT_RELU:
        return np.maximum(x, 0)
    if kind == S.ACT_TANH:
        return np.tanh(x)
    return x / (1 + np.exp(-x))


def run_program(prog, rows, dtype=np.float64):
    rows = np.asarray(rows, dtype=dtype)
    n = len(rows)
    blob = prog.blob.astype(dtype)
    Dp, Hp = S._pad8(prog.D), S._pad8(prog.H)
    bufs = [np.zeros((n, Dp), dtype), np.zeros((n, Dp), dtype), np.zeros((n, Hp), dtype), np.zeros((n, Hp), dtype)]
    bufs[S.BUF_X0][:, : prog.D] = rows
    ld = np.zeros(n, dtype)
    for op in prog.ops:
        typ, src, dst, src_off, K, N, Np, w_off, b_off, flags = [int(v) for v in op[:10]]
        W = blob[w_off : w_off + K * Np].reshape(K, Np)
        b = blob[b_off : b_off + Np]
        a = bufs[src][:, src_off : src_off + K]
        if typ == S.OP_LINEAR:
            if flags & S.FLAG_IN_ACT:
                a = _act(prog.activation, a)
            out = a @ W + b
            if flags & S.FLAG_OUT_ACT:
                out = _act(prog.activation, out)
            if flags & S.FLAG_ACCUM:
                bufs[dst][:, :Np] += out
            else:
                bufs[dst][:, :Np] = out
        elif typ == S.OP_COUPLING_AFFINE:
            xb, d_id, d_tr = int(op[10]), int(op[11]), int(op[12])
            out = a @ W + b
            t = out[:, 0 : 2 * d_tr : 2]
            u = out[:, 1 : 2 * d_tr : 2]
            if flags & S.FLAG_ADDITIVE:
                s = np.ones_like(t)
            elif flags & S.FLAG_SOFTPLUS_SCALE:
                s = np.logaddexp(0, u) + dtype(1e-3)
            else:
                s = 1 / (1 + np.exp(-(u + 2))) + dtype(1e-3)
            x = bufs[xb][:, d_id : d_id + d_tr]
            ls = 0.0 if flags & S.FLAG_NO_LOGDET else np.log(s).sum(1)
            if flags & S.FLAG_INVERSE:
                bufs[dst][:, d_id : d_id + d_tr] = (x - t) / s
                ld -= ls
            else:
                bufs[dst][:, d_id : d_id + d_tr] = x * s + t
                ld += ls
        else:
            raise NotImplementedError(typ)
    return bufs[prog.final_buf][:, : prog.D].copy(), ld + dtype(prog.const_logdet)
