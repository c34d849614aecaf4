"""Config 1 (BASELINE.json): the reference's own FlowSampler / NestedSampler / Model /
live-point code, UNMODIFIED, with our proposal class plugged in (plugin point P3)."""

import numpy as np
import pytest
from conftest import reference_or_skip

pytestmark = [pytest.mark.gpu, pytest.mark.reference]


def make_model():
    from nessai.model import Model

    class Gaussian2D(Model):
        """/root/reference/examples/2d_gaussian.py:28-61"""

        def __init__(self):
            self.names = ["x", "y"]
            self.bounds = {"x": [-10, 10], "y": [-10, 10]}

        def log_prior(self, x):
            log_p = np.log(self.in_bounds(x), dtype="float")
            for n in self.names:
                log_p -= np.log(self.bounds[n][1] - self.bounds[n][0])
            return log_p

        def log_likelihood(self, x):
            log_l = np.zeros(x.size)
            for n in self.names:
                log_l += -0.5 * x[n] ** 2 - 0.5 * np.log(2 * np.pi)
            return log_l

    return Gaussian2D()


def test_flowsampler_runs_with_b200_proposal(tmp_path):
    reference_or_skip()
    from nessai.flowsampler import FlowSampler

    from nessai_b200.nessai_plugin import B200NessaiFlowProposal

    fs = FlowSampler(
        make_model(), output=str(tmp_path), resume=False, seed=1234, nlive=200, plot=False,
        flow_proposal_class=B200NessaiFlowProposal, flow_config=dict(n_blocks=2),
        training_config=dict(max_epochs=50, patience=10), maximum_uninformed=200,
        max_iteration=700, poolsize=2000, checkpointing=False,
    )
    fs.run(plot=False, save=False)
    prop = fs.ns._flow_proposal
    assert isinstance(prop, B200NessaiFlowProposal)
    assert prop.training_count >= 1 and prop.populated_count >= 1
    assert prop._engine is not None  # the fused device loop ran
    assert np.isfinite(fs.ns.log_evidence)
    # analytic log Z = -log(400) = -5.99 for the unit Gaussian in [-10, 10]^2; the run is
    # truncated at max_iteration, so only a loose sanity bound is asserted here
    assert -9.0 < fs.ns.log_evidence < -4.0


def test_flowsampler_with_truncation_rules_on_device(tmp_path):
    """The optional truncation rules (truncation.py:368-429) and the pool likelihood run inside
    the fused device loop when the model offers ``log_likelihood_torch``: the reference's
    sampler, unmodified, never evaluates the likelihood of a pool on the host."""
    reference_or_skip()
    import torch
    from nessai.flowsampler import FlowSampler

    from nessai_b200.nessai_plugin import B200NessaiFlowProposal

    model = make_model()
    calls = {"device_rows": 0}

    def log_likelihood_torch(x):
        calls["device_rows"] += x.shape[0]
        return (-0.5 * x * x - 0.5 * float(np.log(2 * np.pi))).sum(dim=1)

    type(model).log_likelihood_torch = staticmethod(log_likelihood_torch)
    fs = FlowSampler(
        model, output=str(tmp_path), resume=False, seed=4321, nlive=200, plot=False,
        flow_proposal_class=B200NessaiFlowProposal, flow_config=dict(n_blocks=2),
        training_config=dict(max_epochs=50, patience=10), maximum_uninformed=200,
        max_iteration=1000, poolsize=2000, checkpointing=False,
        truncation_methods=["latent_radius", "min_log_q", "likelihood_threshold"],
    )
    fs.run(plot=False, save=False)
    prop = fs.ns._flow_proposal
    assert [r.name for r in prop._truncation_scheme.rules] == ["latent_radius", "min_log_q", "likelihood_threshold"]
    assert prop._engine is not None and prop.populated_count >= 1
    assert calls["device_rows"] > 0 and prop._engine.likelihood is not None
    assert np.isfinite(prop._engine.min_log_q)
    # every pooled sample is above the contour it was drawn for and carries the model's logL
    s = prop.samples
    np.testing.assert_allclose(s["logL"], model.log_likelihood(s), rtol=1e-10)
    assert np.all(s["logL"] > prop._truncation_scheme.rules[2].threshold)
    # truncated at max_iteration: a loose sanity bound around the analytic -5.99
    assert np.isfinite(fs.ns.log_evidence) and -8.0 < fs.ns.log_evidence < -4.5


def test_plugin_matches_reference_flow_numerics(tmp_path):
    """Same weights in the reference FlowModel (CPU, shim) and B200FlowModel."""
    reference_or_skip()
    import torch
    from nessai.flowmodel import FlowModel

    from nessai_b200.flowmodel import B200FlowModel

    cfg = dict(n_inputs=4, n_neurons=8, n_blocks=3, n_layers=2, ftype="realnvp")
    torch.manual_seed(3)
    ref = FlowModel(flow_config=dict(cfg), output=str(tmp_path / "ref"))
    ref.initialise()
    x = np.random.default_rng(0).normal(size=(500, 4))
    ref.train(x, max_epochs=5, plot=False)
    ours = B200FlowModel(flow_config=dict(cfg), output=str(tmp_path / "ours"))
    ours.initialise()
    ours.load_weights(ref.weights_file)  # reference-written model.pt
    z = np.random.default_rng(1).normal(size=(300, 4))
    xr, lr = ref.sample_and_log_prob(z=z)
    xo, lo = ours.sample_and_log_prob(z=z)
    np.testing.assert_allclose(xo, xr, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(lo, lr, rtol=1e-4, atol=1e-4)
    zr, pr = ref.forward_and_log_prob(x)
    zo, po = ours.forward_and_log_prob(x)
    np.testing.assert_allclose(zo, zr, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(po, pr, rtol=1e-4, atol=1e-4)
    # and the other way round: our weights file loads into the reference flow
    ours.save_weights(str(tmp_path / "ours" / "w.pt"))
    ref.load_weights(str(tmp_path / "ours" / "w.pt"))
    np.testing.assert_allclose(ref.sample_and_log_prob(z=z)[1], lr, rtol=1e-6, atol=1e-6)


def test_importance_nested_sampler_runs_on_b200_flows(tmp_path):
    """Config 5 (BASELINE.json): the reference's ImportanceNestedSampler, unmodified,
    with one B200 flow per level (pattern of
    /root/reference/examples/importance_nested_sampler/basic_ins_example.py)."""
    reference_or_skip()
    from nessai.model import Model

    from nessai_b200.importance import B200ImportanceFlowModel
    from nessai_b200.nessai_plugin import B200ImportanceNestedSampler

    class Gaussian4D(Model):
        def __init__(self):
            self.names = [f"x{i}" for i in range(4)]
            self.bounds = {n: [-8.0, 8.0] for n in self.names}

        def log_prior(self, x):
            log_p = np.log(self.in_bounds(x), dtype="float")
            for n in self.names:
                log_p -= np.log(self.bounds[n][1] - self.bounds[n][0])
            return log_p

        def log_likelihood(self, x):
            log_l = np.zeros(x.size)
            for n in self.names:
                log_l += -0.5 * x[n] ** 2 - 0.5 * np.log(2 * np.pi)
            return log_l

        def to_unit_hypercube(self, x):
            x_out = x.copy()
            for n in self.names:
                x_out[n] = (x[n] - self.bounds[n][0]) / (self.bounds[n][1] - self.bounds[n][0])
            return x_out

        def from_unit_hypercube(self, x):
            x_out = x.copy()
            for n in self.names:
                x_out[n] = (self.bounds[n][1] - self.bounds[n][0]) * x[n] + self.bounds[n][0]
            return x_out

    model = Gaussian4D()
    ins = B200ImportanceNestedSampler(
        model, output=str(tmp_path), nlive=500, max_iteration=4, min_samples=100, plot=False,
        checkpointing=False, seed=1234, draw_constant=True, reset_flow=True,
        flow_config=dict(n_blocks=2, n_neurons=16), training_config=dict(max_epochs=30, patience=10),
    )
    ins.nested_sampling_loop()
    assert isinstance(ins.proposal.flow, B200ImportanceFlowModel)
    assert ins.proposal.flow.n_models >= 2
    assert np.isfinite(ins.log_evidence)
    # analytic log Z = -4 log(16) = -11.09; a handful of levels gets within a few units
    assert -14.0 < ins.log_evidence < -8.0


@pytest.mark.parametrize("which", ["b200flowproposal", "b200augmentedflowproposal"])
def test_sampling_resume_with_b200_proposal(tmp_path, which):
    """Checkpoint -> ``FlowSampler(resume=True)`` -> continue, with the B200 proposal classes
    named as the entry-point strings a user would pass
    (/root/reference/tests/test_sampling/test_standard_sampling.py:198-236; the checkpoint pickles
    the proposal, samplers/base.py:346, and the flow model's ``__getstate__`` drops the device
    state, flowmodel/base.py:921-957)."""
    reference_or_skip()
    import os

    from nessai.flowsampler import FlowSampler

    from nessai_b200 import nessai_plugin
    from nessai_b200.flowmodel import B200FlowModel

    cls = dict(b200flowproposal=nessai_plugin.B200NessaiFlowProposal,
               b200augmentedflowproposal=nessai_plugin.B200AugmentedFlowProposal)[which]
    extra = dict(augment_dims=1) if which == "b200augmentedflowproposal" else {}
    output = str(tmp_path / "resume")
    flow_config = dict(n_blocks=2, n_neurons=8)
    fs = FlowSampler(
        make_model(), output=output, resume=True, nlive=100, plot=False, flow_config=flow_config,
        flow_proposal_class=cls, training_config=dict(max_epochs=20, patience=5),
        training_frequency=10, maximum_uninformed=9, checkpoint_on_iteration=True,
        checkpoint_interval=5, seed=1234, max_iteration=11, poolsize=10, **extra,
    )
    fs.run(plot=False, save=False)
    assert os.path.exists(os.path.join(output, "nested_sampler_resume.pkl"))
    assert isinstance(fs.ns._flow_proposal, cls) and fs.ns._flow_proposal.training_count >= 1

    # a new model instance emulates a new run
    fs = FlowSampler(make_model(), output=output, resume=True, flow_config=flow_config,
                     flow_proposal_class=cls, plot=False)
    assert fs.ns.iteration == 11
    prop = fs.ns._flow_proposal
    assert isinstance(prop, cls) and isinstance(prop.flow, B200FlowModel)
    fs.ns.max_iteration = 21
    fs.run(plot=False, save=False)
    assert fs.ns.iteration == 21
    assert os.path.exists(os.path.join(output, "nested_sampler_resume.pkl.old"))
    assert np.isfinite(fs.ns.log_evidence)


def test_complete_run_recovers_the_analytic_evidence(tmp_path):
    """The reference's sampler, unmodified, run TO CONVERGENCE with the B200 proposal and the
    reference's default flow_config on a 4-D unit Gaussian in [-10, 10]^4: log Z = -4 log 20 =
    -11.98 analytically; the nested-sampling error is sqrt(H / nlive) ~ 0.11."""
    reference_or_skip()
    from nessai.flowsampler import FlowSampler
    from nessai.model import Model

    from nessai_b200.nessai_plugin import B200NessaiFlowProposal

    D = 4

    class Gaussian(Model):
        def __init__(self):
            self.names = [f"x{i}" for i in range(D)]
            self.bounds = {n: [-10.0, 10.0] for n in self.names}

        def log_prior(self, x):
            return np.log(self.in_bounds(x), dtype="float") - D * np.log(20.0)

        def log_likelihood(self, x):
            return -0.5 * np.sum(self.unstructured_view(x) ** 2, axis=-1) - 0.5 * D * np.log(2 * np.pi)

    fs = FlowSampler(Gaussian(), output=str(tmp_path), resume=False, seed=2024, nlive=500, plot=False,
                     flow_proposal_class=B200NessaiFlowProposal, checkpointing=False)
    fs.run(plot=False, save=False)
    prop = fs.ns._flow_proposal
    assert prop._engine is not None and prop.training_count >= 2 and prop.populated_count >= 2
    assert prop.flow.model.spec.H == 2 * D and prop.flow.model.spec.net == "resnet"  # the reference's defaults
    assert abs(fs.ns.log_evidence - (-D * np.log(20.0))) < 0.5, fs.ns.log_evidence
    # posterior moments of the unit Gaussian
    post = fs.posterior_samples
    m = np.array([post[n].mean() for n in fs.ns.model.names])
    s = np.array([post[n].std() for n in fs.ns.model.names])
    assert np.all(np.abs(m) < 0.25) and np.all(np.abs(s - 1.0) < 0.2), (m, s)


def test_complete_importance_run_recovers_the_analytic_evidence(tmp_path):
    """Config 5 to convergence: the reference's ImportanceNestedSampler, unmodified, its own stopping
    criterion and default flow_config, one B200 flow per level; 4-D unit Gaussian in [-8, 8]^4:
    log Z = -4 log 16 = -11.09."""
    reference_or_skip()
    from nessai.model import Model

    from nessai_b200.nessai_plugin import B200ImportanceNestedSampler

    class Gaussian4D(Model):
        def __init__(self):
            self.names = [f"x{i}" for i in range(4)]
            self.bounds = {n: [-8.0, 8.0] for n in self.names}

        def log_prior(self, x):
            return np.log(self.in_bounds(x), dtype="float") - 4 * np.log(16.0)

        def log_likelihood(self, x):
            return -0.5 * np.sum(self.unstructured_view(x) ** 2, axis=-1) - 2.0 * np.log(2 * np.pi)

        def to_unit_hypercube(self, x):
            u = x.copy()
            for n in self.names:
                u[n] = (x[n] + 8.0) / 16.0
            return u

        def from_unit_hypercube(self, u):
            x = u.copy()
            for n in self.names:
                x[n] = 16.0 * u[n] - 8.0
            return x

    ins = B200ImportanceNestedSampler(Gaussian4D(), output=str(tmp_path), nlive=1000, plot=False, checkpointing=False,
                                      seed=1234)
    ins.nested_sampling_loop()
    assert ins.proposal.flow.n_models >= 3
    assert abs(ins.log_evidence - (-4 * np.log(16.0))) < 0.3, ins.log_evidence
