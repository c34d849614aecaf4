"""The reference's OWN known answers for this path (SURVEY.md 8c pin table), applied to the
host mirror in ``nessai_b200`` -- no reference import needed, the expected values are the ones
the reference's tests assert (file:line cited per test).  CPU only; the numeric pins that need
the kernels are in the ``-m gpu`` tests."""

import numpy as np
import pytest

from nessai_b200.flowmodel import B200FlowModel
from nessai_b200.proposal import B200FlowProposal, compute_radius


# ---- /root/reference/tests/test_flowmodel/test_flowmodel_base.py:145-183
@pytest.mark.parametrize("data_size, batch_size", [(4010, 1000), (106, 21), (1000, 1000), (2000, 1000)])
def test_check_batch_size(data_size, batch_size):
    out = B200FlowModel.check_batch_size(np.arange(data_size), batch_size)
    assert out >= int(0.1 * batch_size)
    if not data_size % batch_size:
        assert out == batch_size


def test_check_batch_size_min_size():
    """The minimum valid size is 8 but that leaves a final batch of 1, so 7."""
    assert B200FlowModel.check_batch_size(np.arange(25), 10, min_fraction=0.8) == 7


def test_check_batch_size_errors():
    with pytest.raises(RuntimeError, match="Could not find a valid batch size"):
        B200FlowModel.check_batch_size(np.arange(3), 2, min_fraction=1.0)
    with pytest.raises(ValueError, match="Cannot use a batch size of 1"):
        B200FlowModel.check_batch_size(np.arange(2), 1, min_fraction=1.0)


# ---- tests/test_proposal/test_flowproposal/test_flowproposal/test_flowproposal_integration.py:149-181
def test_constant_volume_radius():
    """q = 0.8647 in two dimensions gives a radius of ~2 (4 significant figures)."""
    np.testing.assert_approx_equal(compute_radius(2, 0.8647), 2.0, 4)


class _Model:
    names = ["x", "y"]
    bounds = {"x": [-5.0, 5.0], "y": [-5.0, 5.0]}


def _proposal(**kw):
    return B200FlowProposal(_Model(), rng=np.random.default_rng(0), poolsize=10, **kw)


# ---- tests/test_proposal/test_flowproposal/test_base/test_weights.py:92-102
@pytest.mark.parametrize("acceptance, scale", [(0.0, 10.0), (0.5, 2.0), (0.01, 10.0), (2.0, 1.0)])
def test_update_poolsize_scale(acceptance, scale):
    p = _proposal(max_poolsize_scale=10.0)
    p.update_poolsize_scale(acceptance)
    assert p._poolsize_scale == scale
    assert p.poolsize == int(scale * 10)  # test_base/test_properties.py:19-23


# ---- tests/test_proposal/test_flowproposal/test_base/test_flow.py:107-143
class _StubFlow:
    """What the reference's test mocks: numpy_array_to_tensor and a base log-prob of zero."""

    class model:
        calls = 0

        @classmethod
        def base_distribution_log_prob(cls, z):
            import torch

            cls.calls += 1
            return torch.zeros(z.shape[0])

    @staticmethod
    def numpy_array_to_tensor(x):
        import torch

        return torch.from_numpy(np.asarray(x)).type(torch.get_default_dtype())


def test_latent_log_prob_with_temperature():
    p = _proposal()
    p.flow = _StubFlow()
    _StubFlow.model.calls = 0
    out = p.latent_log_prob(np.array([[2.0, 0.0]]), temperature=4.0)
    assert _StubFlow.model.calls == 1
    np.testing.assert_allclose(out, np.array([-np.log(2.0) * 2]))


@pytest.mark.parametrize("temperature", [None, 1.0])
def test_latent_log_prob_without_temperature_scaling(temperature):
    p = _proposal()
    p.flow = _StubFlow()
    np.testing.assert_allclose(p.latent_log_prob(np.array([[2.0, 0.0]]), temperature=temperature), np.array([0.0]))


# ---- tests/test_proposal/test_flowproposal/test_base/test_configuration.py (latent temperature)
def test_latent_temperature_validation():
    with pytest.raises(TypeError):
        _proposal(latent_temperature="hot")
    with pytest.raises(ValueError):
        _proposal(latent_temperature=0.0)
    assert _proposal(latent_temperature=2).latent_temperature == 2.0


# ---- test_flowproposal_integration.py:204-227: z-score rescale round trip, log_j = -log_j_inv to 1e-15
def test_zscore_rescale_round_trip():
    from nessai_b200.livepoint import numpy_array_to_live_points

    p = _proposal(fallback_reparameterisation="zscore")
    rng = np.random.default_rng(1)
    x = numpy_array_to_live_points(rng.normal(1.0, 3.0, size=(100, 2)), p.names)
    p.check_state(x)
    np.testing.assert_allclose(p.scale, [np.std(x["x"]), np.std(x["y"])], rtol=0, atol=0)
    x_prime, log_j = p.rescale(x)
    for pn in p.prime_parameters:
        np.testing.assert_allclose([x_prime[pn].mean(), x_prime[pn].std()], [0.0, 1.0], atol=1e-12)
    x_back, log_j_inv = p.inverse_rescale(x_prime)
    for n in p.names:
        np.testing.assert_allclose(x_back[n], x[n], rtol=0, atol=1e-14)
    np.testing.assert_allclose(log_j, -log_j_inv, rtol=0, atol=1e-15)


# ---- tests/test_flows/test_flow_utils.py:39-46: silu(x) = x * expit(x), 6 decimals
def test_silu():
    from _program_interp import _act
    from scipy.special import expit

    from nessai_b200 import spec as S
    from oracle.flow_numpy import activation

    x = np.random.default_rng(0).standard_normal(100)
    np.testing.assert_array_almost_equal(activation("silu")(x), x * expit(x))  # the oracle's
    np.testing.assert_array_almost_equal(_act(S.ACT_SILU, x), x * expit(x))  # the program interpreter's


# ---- tests/test_flows/test_flow_utils.py:158-173: BatchNorm reset constants
def test_reset_weights_batch_norm_constants():
    import torch

    from nessai_b200.spec import FlowSpec

    sp = FlowSpec(dict(n_inputs=2, n_neurons=4, n_blocks=2, n_layers=1, ftype="realnvp"))
    torch.manual_seed(0)
    theta, ints = sp.init_state()
    theta = theta + np.float32(0.37)  # "trained" values everywhere, running statistics included
    sp.reset_weights(theta)
    sd = sp.state_dict_numpy(theta, ints)
    constant = np.float32(np.log(np.exp(1 - sp.BN_EPS) - 1))
    seen = 0
    for k, v in sd.items():
        if k.endswith("unconstrained_weight"):
            assert (v == constant).all()
            seen += 1
        elif k.endswith("running_mean") or (k.endswith(".bias") and k.replace(".bias", ".running_mean") in sd):
            assert (v == 0).all()
        elif k.endswith("running_var"):
            assert (v == 1).all()
    assert seen == 2


# ---- tests/test_flows/test_distributions/test_multivariate_normal.py:8-38: the base distribution
# N(0, var I) (var = 1 is the default StandardNormal of every other flow here) against scipy
@pytest.mark.parametrize("dims", [2, 4])
@pytest.mark.parametrize("var", [1, 2, 4])
def test_base_distribution_log_prob(dims, var):
    from scipy import stats

    from oracle.flow_numpy import NumpyFlow
    from oracle.populate_numpy import populate_turn

    class Identity:  # a flow that does nothing: log q is the base density
        base_var = float(var)
        base_log_prob = NumpyFlow.base_log_prob

        @staticmethod
        def inverse(z):
            return z.copy(), np.zeros(len(z))

    z = np.random.default_rng(2).random((1000, dims))
    t = populate_turn(Identity(), z, scale=np.ones(dims), shift=np.zeros(dims), lo=-np.inf, hi=np.inf,
                      log_prior_const=0.0)
    ref = stats.multivariate_normal(mean=np.zeros(dims), cov=var * np.eye(dims)).logpdf(z)
    np.testing.assert_array_almost_equal(t["log_q"], ref)
