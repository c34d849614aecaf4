"""CPU restatement (float64 numpy) of the populate tail for per-parameter maps that are not a
diagonal affine -- TEST INFRASTRUCTURE ONLY: nothing under ``nessai_b200/`` may import this.

Restates the inverse direction of the reference's ``RescaleToBounds``
(/root/reference/src/nessai/reparameterisations/rescale.py:635-660) for the configurations the
device tail covers, and of ``ScaleAndShift`` (rescale.py:263-291) with a pre- or post-rescaling
function, as ``x = h(a x' + b) * scale + shift`` per parameter:

* ``post_rescaling="logit"``: ``h = sigmoid``, ``log|J| += log h + log1p(-h)``
  (utils/rescaling.py:310-330), then ``[0, 1] -> [lo, hi]`` (rescale.py:544-553);
* ``post_rescaling="log"``: ``h = exp``, ``log|J| += x'`` (utils/rescaling.py:385-402);
* boundary inversion (rescale.py:570-590): ``h = |x'|``; a "lower" edge maps ``[0, 1] -> [lo, hi]``,
  an "upper" edge ``1 - |x'|`` first (a negative scale); no detected edge: ``[-1, 1] -> [lo, hi]``;
* ``post_rescaling`` "exp" / "gaussian_cdf" / "inv_gaussian_cdf": ``h`` = log / the normal quantile
  function / the normal CDF (utils/rescaling.py:369-407);
* the same functions as ``pre_rescaling`` ("z-score-logit", "log-z-score",
  "z-score-inv-gaussian-cdf"): ``x = P^-1(scale * x' + shift)``, i.e. ``a = scale``, ``b = shift``;
* ``Angle`` (reparameterisations/angle.py:17-186): the flow sees the Cartesian pair
  ``(r cos(scale * angle), r sin(scale * angle))``; inverse ``r = sqrt(x'^2 + y'^2)``,
  ``angle = atan2(y', x') [% 2 pi] / scale``, ``log|J| -= log r``; ``r`` is either a model parameter or
  an auxiliary one with a ``chi(2)`` prior;
* ``ToCartesian`` (angle.py:189-232): the same Cartesian pair for a NON-periodic parameter mapped to
  ``[0, scale]`` (and mirrored); inverse ``|atan2(y', x') / scale| * (hi - lo) + lo``,
  ``log|J| += log(hi - lo) - log r``;
* ``AnglePair`` (angle.py:235-538): the flow sees the Cartesian triple of two angles and a radius
  (a model parameter or an auxiliary one with a ``chi(3)`` prior), convention "az-zen" or "ra-dec";
* ``Dequantise`` (reparameterisations/discrete.py): ``RescaleToBounds`` whose pre-rescaling inverse
  is ``floor`` with no log-Jacobian;
* ``h = identity``: the diagonal affine of ``oracle/populate_numpy.py``;

followed by ``log_q -= log|J|`` (flowproposal.py:378-383), the prior-bounds check
(flowproposal/base.py:939-959) and ``log_w = log_prior - log_q`` (base.py:1069-1098).

Pinned by ``tests/test_reparam_oracle.py`` against the reference's own
``FlowProposal.inverse_rescale`` on the CPU.
"""

from __future__ import annotations

import numpy as np

IDENTITY, SIGMOID, ABS, EXP, LOG, NORMAL_CDF, NORMAL_QUANTILE = range(7)
# pair kinds: functions of two flow features, reparameterisations/angle.py:149-181
ANGLE, ANGLE_MOD, RADIUS, RADIUS_CHI = 7, 8, 9, 10
FLOOR = 11  # single feature (Dequantise)
ANGLE_ABS = 12  # pair (ToCartesian)
ZENITH, DECLINATION, RADIUS3, RADIUS3_CHI = 13, 14, 15, 16  # triples (AnglePair)
GAUSS_AUX = 17  # single feature: an augment parameter with its N(0, 1) prior (proposal/augmented.py:162-178)
FLOOR_AFTER = 0x100  # flag on a single-feature kind: floor of the final value ("dequantise-logit")
KIND_MASK = 0xFF


def is_multi(kind):
    """Kinds that read two or three flow features."""
    kind = np.asarray(kind) & KIND_MASK
    return (kind >= ANGLE) & (kind != FLOOR) & (kind != GAUSS_AUX)



def inverse_maps(xp, kind, scale, shift, pre_scale=None, pre_shift=None, src=None, return_log_prior=False):
    """``x' (n, D) -> (x (n, D), log|J| (n,))`` for ``x = h(a x' + b) * scale + shift`` with the
    per-parameter ``kind`` of ``h``, ``a = pre_scale`` (default 1) and ``b = pre_shift`` (0).
    ``src`` (``(D, 3)`` ints -- ``(D, 2)`` is padded --, default ``[d, d, d]``): the flow feature(s)
    output slot ``d`` reads; the pair kinds read two -- ``ANGLE``: ``atan2(u1, u0) * scale + shift`` (``ANGLE_MOD``: modulo
    ``2 pi`` first), no log-Jacobian for the constant factor (angle.py:120-128,157-170);
    ``RADIUS``: ``sqrt(u0^2 + u1^2)``, ``log|J| -= log r`` (angle.py:172); ``RADIUS_CHI``: the same for
    an auxiliary radius, whose ``chi(2)`` prior ``log r - r^2 / 2`` (angle.py:183-185) is returned
    as a third array with ``return_log_prior``.  ``ANGLE_ABS`` (ToCartesian, angle.py:221-231):
    ``|atan2(u1, u0) * a| * scale + shift``, ``log|J| += log|scale|``.  Triples (AnglePair,
    angle.py:418-489): ``ZENITH`` ``atan2(sqrt(u0^2 + u1^2), u2)`` with ``log|J| -= log sin``;
    ``DECLINATION`` ``atan2(u2, sqrt(u0^2 + u1^2))`` with ``log|J| -= log cos``; ``RADIUS3(_CHI)``
    ``sqrt(u0^2 + u1^2 + u2^2)`` with ``log|J| -= 2 log r`` (and the ``chi(3)`` prior, :529-537).
    ``FLOOR`` (Dequantise): ``floor(u)``, no log-Jacobian."""
    from scipy.special import erfc, erfcinv

    xp = np.asarray(xp, dtype=np.float64)
    floor_after = (np.asarray(kind) & FLOOR_AFTER) != 0
    kind = np.asarray(kind) & KIND_MASK
    D = xp.shape[1]
    a = np.ones(D) if pre_scale is None else np.asarray(pre_scale, dtype=np.float64)
    b = np.zeros(D) if pre_shift is None else np.asarray(pre_shift, dtype=np.float64)
    src = np.stack([np.arange(D)] * 3, axis=1) if src is None else np.asarray(src).reshape(D, -1)
    if src.shape[1] == 2:
        src = np.concatenate([src, src[:, :1]], axis=1)
    x = np.empty_like(xp)
    single = ~is_multi(kind)
    log_j = np.full(xp.shape[0], float(np.sum(np.log(np.abs(np.asarray(scale)[single])))
                                       + np.sum(np.log(np.abs(a[single])))
                                       + np.sum(np.log(np.abs(np.asarray(scale)[kind == ANGLE_ABS])))))
    log_p = np.zeros(xp.shape[0])
    with np.errstate(all="ignore"):
        for d in range(D):
            if is_multi(kind[d]):
                u0, u1, u2 = xp[:, src[d, 0]], xp[:, src[d, 1]], xp[:, src[d, 2]]
                if kind[d] == ANGLE_ABS:
                    h = np.abs(np.arctan2(u1, u0) * a[d])
                elif kind[d] == ZENITH:
                    h = np.arctan2(np.sqrt(u0**2 + u1**2), u2)
                    log_j = log_j - np.log(np.sin(h))
                elif kind[d] == DECLINATION:
                    h = np.arctan2(u2, np.sqrt(u0**2 + u1**2))
                    log_j = log_j - np.log(np.cos(h))
                elif kind[d] in (RADIUS3, RADIUS3_CHI):
                    h = np.sqrt(u0**2 + u1**2 + u2**2)
                    log_j = log_j - 2.0 * np.log(h)
                    if kind[d] == RADIUS3_CHI:
                        log_p = log_p + 0.5 * np.log(2.0 / np.pi) + 2.0 * np.log(h) - 0.5 * h**2
                elif kind[d] in (ANGLE, ANGLE_MOD):
                    h = np.arctan2(u1, u0)
                    if kind[d] == ANGLE_MOD:
                        h = h % (2.0 * np.pi)
                elif kind[d] in (RADIUS, RADIUS_CHI):
                    h = np.sqrt(u0**2 + u1**2)
                    log_j = log_j - np.log(h)
                    if kind[d] == RADIUS_CHI:
                        log_p = log_p + np.log(h) - 0.5 * h**2
                else:
                    raise ValueError(f"unknown kind {kind[d]}")
                x[:, d] = h * scale[d] + shift[d]
                continue
            u = a[d] * xp[:, src[d, 0]] + b[d]
            if kind[d] == SIGMOID:  # utils/rescaling.py:310-330
                h = 1.0 / (1.0 + np.exp(-u))
                log_j = log_j + np.log(h) + np.log1p(-h)
            elif kind[d] == ABS:  # rescale.py:570-590
                h = np.abs(u)
            elif kind[d] == EXP:  # utils/rescaling.py:385-393
                h = np.exp(u)
                log_j = log_j + u
            elif kind[d] == LOG:  # utils/rescaling.py:369-383
                h = np.log(u)
                log_j = log_j - h
            elif kind[d] == NORMAL_CDF:  # utils/rescaling.py:396-400
                h = 0.5 * erfc(-u / np.sqrt(2.0))
                log_j = log_j - 0.5 * np.log(2 * np.pi) - 0.5 * u**2
            elif kind[d] == NORMAL_QUANTILE:  # utils/rescaling.py:403-407
                h = -np.sqrt(2.0) * erfcinv(2.0 * u)
                log_j = log_j + 0.5 * np.log(2 * np.pi) + 0.5 * h**2
            elif kind[d] == FLOOR:  # reparameterisations/discrete.py:77-78
                h = np.floor(u)
            elif kind[d] == GAUSS_AUX:  # proposal/augmented.py:150-178
                h = u
                log_p = log_p - 0.5 * u**2 - 0.5 * np.log(2 * np.pi)
            elif kind[d] == IDENTITY:
                h = u
            else:
                raise ValueError(f"unknown kind {kind[d]}")
            x[:, d] = h * scale[d] + shift[d]
            if floor_after[d]:  # Dequantise with a post-rescaling: floor last, no log-Jacobian
                x[:, d] = np.floor(x[:, d])
    if return_log_prior:
        return x, log_j, log_p
    return x, log_j


def tail_rows(xp, logq_flow, *, kind, scale, shift, lo, hi, log_prior_const, min_log_q=None, pre_scale=None,
              pre_shift=None, src=None):
    """The tail of one turn: ``(x, log_q, log_w, valid)``; ``logq_flow`` is the flow's own
    ``log q`` (NaN where the row was dropped before: radius truncation, non-finite)."""
    x, log_j, log_p = inverse_maps(xp, kind, scale, shift, pre_scale, pre_shift, src, return_log_prior=True)
    with np.errstate(all="ignore"):
        log_q = np.asarray(logq_flow, dtype=np.float64) - log_j
        log_w = log_prior_const + log_p - log_q
        valid = np.isfinite(log_q) & np.isfinite(log_w) & ~np.any((x < lo) | (x > hi), axis=1)
        if min_log_q is not None:
            valid &= log_q > min_log_q
    return x, np.where(valid, log_q, np.nan), np.where(valid, log_w, np.nan), valid
