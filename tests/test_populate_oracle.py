"""Pins ``oracle/populate_numpy.py`` (the float64 restatement of FlowProposal.populate's turn,
rejection step and loop) against the reference's OWN ``populate`` run on the CPU: the
reference's latent draws and uniforms are recorded and replayed through the oracle."""

import json
import os

import numpy as np
import pytest
from conftest import GOLDEN, reference_or_skip

pytestmark = pytest.mark.reference


class RecordingRng:
    """numpy Generator proxy that keeps every ``random(n)`` block it hands out."""

    def __init__(self, rng):
        self._rng, self.blocks = rng, []

    def random(self, *a, **k):
        u = self._rng.random(*a, **k)
        self.blocks.append(np.array(u, copy=True))
        return u

    def __getattr__(self, name):
        return getattr(self._rng, name)


def reference_proposal(tmp_path, drawsize, **kw):
    import torch
    from nessai.livepoint import numpy_array_to_live_points
    from nessai.model import Model
    from nessai.proposal import FlowProposal

    g = np.load(os.path.join(GOLDEN, "c2_realnvp_mlp.npz"))
    cfg = json.loads(str(g["flow_config"]))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    D = cfg["n_inputs"]

    class Box(Model):
        def __init__(self):
            self.names = [f"x{i}" for i in range(D)]
            self.bounds = {n: [-4.0, 4.0] for n in self.names}

        def log_prior(self, x):
            return np.log(self.in_bounds(x), dtype="float") - D * np.log(8.0)

        def log_likelihood(self, x):
            return -0.5 * np.sum(self.unstructured_view(x) ** 2, axis=-1)

    model = Box()
    rng = RecordingRng(np.random.default_rng(17))
    model.set_rng(rng)
    torch.manual_seed(17)
    prop = FlowProposal(model, rng=rng, flow_config=dict(cfg), output=str(tmp_path), poolsize=drawsize,
                        drawsize=drawsize, plot=False, fallback_reparameterisation="zscore", **kw)
    prop.initialise()
    live = numpy_array_to_live_points(1.5 * np.random.default_rng(5).standard_normal((500, D)) + 0.3, model.names)
    live["logL"] = model.log_likelihood(live)
    prop.check_state(live)
    prop.flow.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    prop.flow.model.eval()
    zs = []
    draw = prop.sample_latent_distribution

    def recording_draw(n):
        z = draw(n)
        zs.append(np.array(z, dtype=np.float64, copy=True))
        return z

    prop.sample_latent_distribution = recording_draw
    return prop, model, live, sd, cfg, zs, rng


def diagonal_rescale(prop, D):
    """(scale, shift) of x = x' * scale + shift, read off the reference's inverse_rescale."""
    from nessai.livepoint import empty_structured_array

    probe = empty_structured_array(2, names=prop.prime_parameters)
    for p in prop.prime_parameters:
        probe[p] = [0.0, 1.0]
    x, log_j = prop.inverse_rescale(probe)
    a = np.stack([x[n] for n in prop.model.names], axis=-1)
    scale, shift = a[1] - a[0], a[0]
    np.testing.assert_allclose(log_j, np.sum(np.log(np.abs(scale))), rtol=1e-12)
    return scale, shift


@pytest.mark.parametrize("rules", [None, ["latent_radius", "min_log_q"], ["latent_radius", "likelihood_threshold"]])
def test_oracle_turn_and_loop_match_reference_populate(tmp_path, rules):
    reference_or_skip()
    from oracle.flow_numpy import NumpyFlow
    from oracle.populate_numpy import populate_loop, populate_turn, rejection_step

    drawsize, n_samples = 4000, 500
    kw = {} if rules is None else dict(truncation_methods=rules)
    prop, model, live, sd, cfg, zs, rng = reference_proposal(tmp_path, drawsize, **kw)
    D = cfg["n_inputs"]
    worst = live[np.argsort(live["logL"])[len(live) // 3]]
    prop.training_data = live
    n_blocks0 = len(rng.blocks)
    prop.populate(worst, n_samples=n_samples, plot=False, max_samples=40 * drawsize)
    us = rng.blocks[n_blocks0:]
    # the reference consumed one block of uniforms per turn (+ nothing else from rng.random)
    assert len(us) == len(zs) and len(zs) >= 2
    scale, shift = diagonal_rescale(prop, D)
    nf = NumpyFlow(sd, ftype="realnvp", net="mlp", hidden_features=cfg["n_neurons"])
    scheme = {r.name: r for r in prop._truncation_scheme.rules}
    turn_kw = dict(
        scale=scale, shift=shift, lo=-4.0, hi=4.0, log_prior_const=-D * np.log(8.0),
        r_max=scheme["latent_radius"].threshold,
        min_log_q=scheme["min_log_q"].min_log_q if "min_log_q" in scheme else None,
        log_likelihood=(lambda x: -0.5 * np.sum(x**2, axis=-1)) if "likelihood_threshold" in scheme else None,
        log_l_threshold=scheme["likelihood_threshold"].threshold if "likelihood_threshold" in scheme else None,
    )
    # ---- turn by turn: same valid rows, same accepted rows (up to fp32-level ties), same x
    kept, n_acc, ambiguous = [], 0, 0
    for z, u in zip(zs, us):
        t = populate_turn(nf, z, **turn_kw)
        assert int(t["valid"].sum()) == len(u)  # the reference drew one uniform per surviving row
        uu = np.full(len(z), np.nan)
        uu[t["valid"]] = u
        accept, margin = rejection_step(t["log_w"], uu)
        ambiguous += int((np.abs(margin[t["valid"]]) < 1e-4).sum())
        m = min(n_samples - n_acc, int(accept.sum()))
        kept.append(t["x"][accept][:m])
        n_acc += int(accept.sum())
    mine = np.concatenate(kept)
    ref = np.stack([prop.samples[n] for n in model.names], axis=-1)
    assert abs(len(mine) - len(ref)) <= ambiguous
    assert ambiguous <= 1  # |margin| < 1e-4 is a ~1e-5-probability event per row
    if ambiguous == 0:
        np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=1e-4)  # the reference's flow is fp32
        assert prop.population_acceptance == n_acc / (len(zs) * drawsize)
        # ---- the loop itself, driven by the recorded draws: same stop condition, same pool
        zi, ui = iter(zs), iter(us)

        def draw_u(m):
            u = next(ui)
            assert len(u) == m
            return u

        x, n_proposed, n_accepted = populate_loop(nf, lambda n: next(zi), draw_u, n_samples, drawsize,
                                                  max_samples=40 * drawsize, **turn_kw)
        assert n_proposed == len(zs) * drawsize and n_accepted == n_acc
        np.testing.assert_array_equal(x, mine)
        assert next(zi, None) is None  # the oracle loop stopped exactly where the reference did
    if "likelihood_threshold" in scheme:
        assert np.all(prop.samples["logL"] > scheme["likelihood_threshold"].threshold)


def test_oracle_accumulate_loop_matches_reference_populate(tmp_path):
    """``accumulate_weights=True`` (flowproposal.py:471-490,504-512): the oracle loop, fed the
    reference's recorded latent draws and uniform blocks, consumes them at the same points,
    stops on the same turn and keeps the same pool."""
    reference_or_skip()
    from oracle.flow_numpy import NumpyFlow
    from oracle.populate_numpy import populate_loop_accumulate

    drawsize, n_samples = 4000, 300
    prop, model, live, sd, cfg, zs, rng = reference_proposal(tmp_path, drawsize, accumulate_weights=True)
    D = cfg["n_inputs"]
    worst = live[np.argsort(live["logL"])[len(live) // 3]]
    prop.training_data = live
    n_blocks0 = len(rng.blocks)
    prop.populate(worst, n_samples=n_samples, plot=False, max_samples=40 * drawsize)
    us = rng.blocks[n_blocks0:]
    assert len(zs) >= 2 and 1 <= len(us) <= len(zs)
    scale, shift = diagonal_rescale(prop, D)
    nf = NumpyFlow(sd, ftype="realnvp", net="mlp", hidden_features=cfg["n_neurons"])
    scheme = {r.name: r for r in prop._truncation_scheme.rules}
    zi, ui = iter(zs), iter(us)

    def draw_u(m):
        u = next(ui)
        assert len(u) == m  # one uniform per accumulated surviving row, as the reference drew
        return u

    x, n_proposed, n_accepted, log_n_expected = populate_loop_accumulate(
        nf, lambda n: next(zi), draw_u, n_samples, drawsize, max_samples=40 * drawsize,
        scale=scale, shift=shift, lo=-4.0, hi=4.0, log_prior_const=-D * np.log(8.0),
        r_max=scheme["latent_radius"].threshold)
    assert next(zi, None) is None and next(ui, None) is None  # same number of turns and of rejection steps
    assert n_proposed == len(zs) * drawsize
    ref = np.stack([prop.samples[n] for n in model.names], axis=-1)
    # the reference's flow is fp32: a row whose margin is at rounding level may flip
    assert abs(len(x) - len(ref)) <= 1 and abs(n_accepted / n_proposed - prop.population_acceptance) <= 1 / n_proposed
    if len(x) == len(ref):
        np.testing.assert_allclose(x, ref, rtol=1e-4, atol=1e-4)
    assert log_n_expected >= np.log(n_samples)


def test_accumulate_control_replays_reference_loop(tmp_path):
    """``nessai_b200.proposal.AccumulateControl`` (the host-side loop control of the device
    accumulate path) takes the reference's decisions: fed the recorded draws it runs the
    rejection step on the same turns, stops on the same turn and repeats the last rejection
    step exactly when the reference does."""
    reference_or_skip()
    from nessai_b200.proposal import AccumulateControl
    from oracle.flow_numpy import NumpyFlow
    from oracle.populate_numpy import populate_turn

    for drawsize, n_samples, max_samples in [(4000, 300, 160000), (3000, 5000, 7000)]:
        prop, model, live, sd, cfg, zs, rng = reference_proposal(tmp_path, drawsize, accumulate_weights=True)
        D = cfg["n_inputs"]
        prop.training_data = live
        n_blocks0 = len(rng.blocks)
        prop.populate(live[0], n_samples=n_samples, plot=False, max_samples=max_samples)
        us = rng.blocks[n_blocks0:]
        scale, shift = diagonal_rescale(prop, D)
        nf = NumpyFlow(sd, ftype="realnvp", net="mlp", hidden_features=cfg["n_neurons"])
        radius = {r.name: r for r in prop._truncation_scheme.rules}["latent_radius"].threshold
        zi, ui = iter(zs), iter(us)
        ctl = AccumulateControl(n_samples, max_samples)
        lws, const = [], -np.inf

        def reject():
            lw = np.concatenate(lws)
            u = next(ui)
            assert len(u) == len(lw)
            ctl.rejected(int(((lw - const) > np.log(u)).sum()))

        while ctl.go_on():
            t = populate_turn(nf, next(zi), scale=scale, shift=shift, lo=-4.0, hi=4.0,
                              log_prior_const=-D * np.log(8.0), r_max=radius)
            v = t["valid"]
            if v.any():
                lws.append(t["log_w"][v])
                const = max(const, float(t["log_w"][v].max()))
            n_expected = float(np.sum(np.exp(np.concatenate(lws) - const))) if lws else 0.0
            if ctl.turn_drawn(drawsize, bool(v.any()), n_expected):
                reject()
            ctl.end_turn()
        if ctl.stale:
            reject()
        assert next(zi, None) is None and next(ui, None) is None
        assert ctl.n_proposed == len(zs) * drawsize
        assert abs(ctl.n_accepted - round(prop.population_acceptance * ctl.n_proposed)) <= 1
        if max_samples == 7000:  # ended by max_samples (3 turns of 3000), pool short of n_samples
            assert ctl.stopped_on_max_samples and len(zs) == 3 and prop.samples.size < n_samples


def test_accumulate_control_edge_cases():
    from nessai_b200.proposal import AccumulateControl

    c = AccumulateControl(100, 1000)
    assert c.go_on()
    assert not c.turn_drawn(400, False, 0.0) and not c.stale  # nothing survived: only max_samples is checked
    c.end_turn()
    assert c.go_on() and c.n_proposed == 400
    assert not c.turn_drawn(400, True, 99.9) and c.stale  # expected pool still short
    c.end_turn()
    assert c.turn_drawn(400, True, 100.0)  # log(100) >= log(100)
    c.rejected(97)
    c.end_turn()
    assert not c.stale and c.stopped_on_max_samples and not c.go_on()  # 1200 > 1000
    c = AccumulateControl(10, 10**6)
    assert c.turn_drawn(50, True, 20.0)
    c.rejected(12)
    c.end_turn()
    assert not c.go_on() and not c.stale and c.n_accepted == 12
