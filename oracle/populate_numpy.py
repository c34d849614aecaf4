"""CPU restatement (float64 numpy) of ONE turn of ``FlowProposal.populate`` and of the loop
around it -- TEST INFRASTRUCTURE ONLY: nothing under ``nessai_b200/`` may import this module.

What is restated (reference file:line):

* the turn, /root/reference/src/nessai/proposal/flowproposal/flowproposal.py:431-469 --
  latent draw ``z`` (supplied by the caller: the reference's stream is torch's CPU ``randn``,
  ours is Philox) -> ``LatentRadiusTruncation.apply_latent`` (truncation.py:352-365) ->
  ``backward_pass`` (flowproposal.py:345-389): ``x', log_j = flow.inverse(z)``,
  ``log_q = latent_log_prob(z) - log_j`` (base.py:401-414, latent temperature), drop
  non-finite rows, diagonal inverse rescale ``x = x' * scale + shift``,
  ``log_q -= sum log|scale|`` (reparameterisations/rescale.py:263-291), prior bounds
  (base.py:939-959, model.py:497-518) -> ``MinLogQTruncation.apply_after_backward``
  (truncation.py:388-394) -> ``LikelihoodThresholdTruncation.apply_after_likelihood``
  (truncation.py:422-429) -> ``log_w = log_prior - log_q`` (base.py:1069-1098);
* the rejection step, flowproposal.py:491-502: ``log_w -= max``, ``accept = log_w > log u``
  (``u`` supplied), the first ``n_samples - n_accepted`` accepted rows kept in draw order;
* the loop and its stop conditions, flowproposal.py:431,436-439,496-502;
* the ``accumulate_weights`` variant of the loop and of the rejection step,
  flowproposal.py:414-417,471-490,504-512: every turn's surviving rows and weights are kept,
  ``log_constant`` is the running maximum, the expected pool size is
  ``exp(logsumexp(log_weights - log_constant))`` and the rejection step runs over ALL rows so
  far (fresh uniforms each time) only once that reaches ``n_samples``.

Pinned by ``tests/test_populate_oracle.py`` against the reference's own ``populate`` run on
the CPU with its latent draws and uniforms recorded.
"""

from __future__ import annotations

import numpy as np


def populate_turn(flow, z, *, scale, shift, lo, hi, log_prior_const, r_max=0.0, sqrt_t=1.0, min_log_q=None,
                  log_likelihood=None, log_l_threshold=None):
    """One turn up to the weights.  ``flow``: ``oracle.flow_numpy.NumpyFlow``; ``z``: ``(n, D)``
    latent draws from the flow's base distribution, BEFORE the temperature scaling.  Returns a dict of per-row arrays over all
    ``n`` rows: ``valid`` (bool), ``x``, ``log_q``, ``log_w`` (NaN where not valid), ``log_l``."""
    z = np.asarray(z, dtype=np.float64) * float(sqrt_t)
    n, D = z.shape
    valid = np.ones(n, dtype=bool)
    if r_max and r_max > 0:
        valid &= np.sqrt(np.sum(z * z, axis=1)) <= r_max  # truncation.py:358-365
    with np.errstate(all="ignore"):
        xp, log_j = flow.inverse(z)
        # base.py:401-414: log N(z / sqrt T) - D log sqrt T
        # (base: N(0, I), or N(0, var I) for flow_config["distribution"] = "mvn", flows/distributions.py:45-56;
        # z0 = z / sqrt T is then a draw from THAT distribution)
        base = flow.base_log_prob(z / sqrt_t) - D * np.log(sqrt_t)
        log_q = base - log_j
        valid &= np.isfinite(log_q)  # flowproposal.py:366-368
        x = xp * scale + shift
        log_q = log_q - np.sum(np.log(np.abs(scale)))
        valid &= ~np.any((x < lo) | (x > hi), axis=1)  # base.py:939-959
        if min_log_q is not None:
            valid &= log_q > min_log_q
        log_l = np.full(n, np.nan)
        if log_likelihood is not None:
            log_l = np.asarray(log_likelihood(x), dtype=np.float64)
            if log_l_threshold is not None:
                valid &= log_l > log_l_threshold
    log_w = np.where(valid, log_prior_const - log_q, np.nan)
    return dict(valid=valid, x=x, log_q=np.where(valid, log_q, np.nan), log_w=log_w, log_l=log_l)


def rejection_step(log_w, u):
    """flowproposal.py:491-494 over the valid rows: ``(accept, margin)`` with
    ``margin = (log_w - max) - log u`` (rows with ``|margin|`` at rounding level are the only
    ones on which two correct implementations may disagree)."""
    valid = ~np.isnan(log_w)
    accept = np.zeros(len(log_w), dtype=bool)
    margin = np.full(len(log_w), np.nan)
    if valid.any():
        margin[valid] = (log_w[valid] - log_w[valid].max()) - np.log(u[valid])
        accept[valid] = margin[valid] > 0
    return accept, margin


def populate_loop(flow, draw_z, draw_u, n_samples, drawsize, max_samples=1_000_000, **turn_kwargs):
    """flowproposal.py:425-502 (no weight accumulation).  ``draw_z(n) -> (n, D)`` and
    ``draw_u(m) -> (m,)`` supply the latent draws of a turn and the uniforms of its VALID rows
    (the reference draws ``rng.random(len(log_w))`` after the truncations).  Returns
    ``(x_accepted[: n_samples], n_proposed, n_accepted)``."""
    out, n_proposed, n_accepted = [], 0, 0
    while n_accepted < n_samples:
        t = populate_turn(flow, draw_z(drawsize), **turn_kwargs)
        n_proposed += drawsize
        if t["valid"].any():
            u = np.full(drawsize, np.nan)
            u[t["valid"]] = draw_u(int(t["valid"].sum()))
            accept, _ = rejection_step(t["log_w"], u)
            m = min(n_samples - n_accepted, int(accept.sum()))
            out.append(t["x"][accept][:m])
            n_accepted += int(accept.sum())
        if n_proposed > max_samples:
            break
    x = np.concatenate(out) if out else np.empty((0, 0))
    return x, n_proposed, n_accepted


def populate_loop_accumulate(flow, draw_z, draw_u, n_samples, drawsize, max_samples=1_000_000, on_valid=None,
                             **turn_kwargs):
    """flowproposal.py:414-417,431-490,504-512 (``accumulate_weights=True``).  ``draw_u(m)`` is
    called exactly where the reference calls ``rng.random(len(log_weights))``: with ``m`` = the
    number of valid rows accumulated so far, every time the expected pool size reaches
    ``n_samples``, and once more after the loop if the last rejection step is stale.
    ``on_valid(mask)`` (optional) is told every turn's surviving-row mask.  Returns
    ``(x_accepted[: n_samples], n_proposed, n_accepted, log_n_expected)``."""
    xs, lws = [], []
    log_constant, log_n_expected = -np.inf, -np.inf
    n_proposed, n_accepted, accept = 0, 0, None
    log_n = np.log(n_samples)
    total = 0
    while n_accepted < n_samples:
        t = populate_turn(flow, draw_z(drawsize), **turn_kwargs)
        n_proposed += drawsize
        v = t["valid"]
        if on_valid is not None:
            on_valid(v)
        if not v.any():  # flowproposal.py:436-439,449-453: nothing survived the truncations
            if n_proposed > max_samples:
                break
            continue
        xs.append(t["x"][v])
        lws.append(t["log_w"][v])
        total += int(v.sum())
        log_constant = max(float(np.max(t["log_w"][v])), log_constant)
        lw = np.concatenate(lws)
        d = lw - log_constant
        log_n_expected = float(np.log(np.sum(np.exp(d))))  # scipy logsumexp of values <= 0
        if log_n_expected >= log_n:
            accept = (lw - log_constant) > np.log(draw_u(total))
            n_accepted = int(accept.sum())
        if n_proposed > max_samples:
            break
    if not xs:
        return np.empty((0, 0)), n_proposed, 0, log_n_expected
    lw = np.concatenate(lws)
    if accept is None or len(accept) != total:
        accept = (lw - log_constant) > np.log(draw_u(total))
    n_accepted = int(accept.sum())
    x = np.concatenate(xs)[accept][:n_samples]
    return x, n_proposed, n_accepted, log_n_expected
