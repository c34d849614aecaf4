"""CPU restatement (float64 numpy) of the populate tail for per-parameter maps that are not a
diagonal affine -- TEST INFRASTRUCTURE ONLY: nothing under ``nessai_b200/`` may import this.

Restates the inverse direction of the reference's ``RescaleToBounds``
(/root/reference/src/nessai/reparameterisations/rescale.py:635-660) for the configurations the
device tail covers, as ``x = h(x') * scale + shift`` per parameter:

* ``post_rescaling="logit"``: ``h = sigmoid``, ``log|J| += log h + log1p(-h)``
  (utils/rescaling.py:310-330), then ``[0, 1] -> [lo, hi]`` (rescale.py:544-553);
* ``post_rescaling="log"``: ``h = exp``, ``log|J| += x'`` (utils/rescaling.py:385-402);
* boundary inversion (rescale.py:570-590): ``h = |x'|``; a "lower" edge maps ``[0, 1] -> [lo, hi]``,
  an "upper" edge ``1 - |x'|`` first (a negative scale); no detected edge: ``[-1, 1] -> [lo, hi]``;
* ``h = identity``: the diagonal affine of ``oracle/populate_numpy.py``;

followed by ``log_q -= log|J|`` (flowproposal.py:378-383), the prior-bounds check
(flowproposal/base.py:939-959) and ``log_w = log_prior - log_q`` (base.py:1069-1098).

Pinned by ``tests/test_reparam_oracle.py`` against the reference's own
``FlowProposal.inverse_rescale`` on the CPU.
"""

from __future__ import annotations

import numpy as np

IDENTITY, SIGMOID, ABS, EXP = 0, 1, 2, 3


def inverse_maps(xp, kind, scale, shift):
    """``x' (n, D) -> (x (n, D), log|J| (n,))`` for the per-parameter ``kind / scale / shift``."""
    xp = np.asarray(xp, dtype=np.float64)
    kind = np.asarray(kind)
    x = np.empty_like(xp)
    log_j = np.full(xp.shape[0], float(np.sum(np.log(np.abs(scale)))))
    with np.errstate(all="ignore"):
        for d in range(xp.shape[1]):
            v = xp[:, d]
            if kind[d] == SIGMOID:
                h = 1.0 / (1.0 + np.exp(-v))
                log_j = log_j + np.log(h) + np.log1p(-h)
            elif kind[d] == ABS:
                h = np.abs(v)
            elif kind[d] == EXP:
                h = np.exp(v)
                log_j = log_j + v
            elif kind[d] == IDENTITY:
                h = v
            else:
                raise ValueError(f"unknown kind {kind[d]}")
            x[:, d] = h * scale[d] + shift[d]
    return x, log_j


def tail_rows(xp, logq_flow, *, kind, scale, shift, lo, hi, log_prior_const, min_log_q=None):
    """The tail of one turn: ``(x, log_q, log_w, valid)``; ``logq_flow`` is the flow's own
    ``log q`` (NaN where the row was dropped before: radius truncation, non-finite)."""
    x, log_j = inverse_maps(xp, kind, scale, shift)
    with np.errstate(all="ignore"):
        log_q = np.asarray(logq_flow, dtype=np.float64) - log_j
        valid = np.isfinite(log_q) & ~np.any((x < lo) | (x > hi), axis=1)
        if min_log_q is not None:
            valid &= log_q > min_log_q
    log_w = np.where(valid, log_prior_const - log_q, np.nan)
    return x, np.where(valid, log_q, np.nan), log_w, valid
