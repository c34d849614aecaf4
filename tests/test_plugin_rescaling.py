"""``nessai_plugin.diagonal_rescaling`` against the reference's own reparameterisation objects
(CPU): whenever it returns ``(scale, shift)`` the reference's ``inverse_rescale`` /
``rescale`` must be exactly that diagonal affine with ``log|J| = sum log|scale|``; anything
that is not a diagonal affine must be refused (the plugin then keeps the host loop)."""

import numpy as np
import pytest
from conftest import reference_or_skip

pytestmark = pytest.mark.reference

D = 4


def make_proposal(tmp_path, **kw):
    from nessai.livepoint import numpy_array_to_live_points
    from nessai.model import Model
    from nessai.proposal import FlowProposal

    class Box(Model):
        def __init__(self):
            self.names = [f"x{i}" for i in range(D)]
            self.bounds = {n: [-4.0 - i, 6.0 + 2 * i] for i, n in enumerate(self.names)}

        def log_prior(self, x):
            return np.log(self.in_bounds(x), dtype="float")

        def log_likelihood(self, x):
            return -0.5 * np.sum(self.unstructured_view(x) ** 2, axis=-1)

    model = Box()
    rng = np.random.default_rng(3)
    model.set_rng(rng)
    prop = FlowProposal(model, rng=rng, flow_config=dict(n_blocks=2, n_neurons=8), output=str(tmp_path),
                        poolsize=100, plot=False, **kw)
    prop.initialise()
    live = numpy_array_to_live_points(1.3 * rng.standard_normal((300, D)) + 0.4, model.names)
    prop.check_state(live)  # what train() does before rescaling the data (base.py:884-893)
    return prop, model, live


CASES = {
    "zscore": dict(fallback_reparameterisation="zscore"),
    "null": dict(fallback_reparameterisation=None),
    "rescaletobounds_updated": dict(fallback_reparameterisation="rescaletobounds"),
    "rescaletobounds_fixed_offset": dict(
        reparameterisations={"rescaletobounds": dict(parameters=[f"x{i}" for i in range(D)], update_bounds=False,
                                                     offset=True, rescale_bounds=[0.0, 1.0])}),
    # boundary inversion whose edge detection found no edge (the data sit away from the bounds):
    # the reference then rescales to [-1, 1] (rescale.py:609-617), a diagonal affine
    "inversion_no_edge": dict(reparameterisations={"inversion": dict(parameters=[f"x{i}" for i in range(D)])}),
    "mixed": dict(reparameterisations={"x0": "default", "x1": "z-score", "x2": "null",
                                       "x3": {"reparameterisation": "scale", "scale": 2.5}}),
}


@pytest.mark.parametrize("case", list(CASES))
def test_diagonal_rescaling_is_the_reference_map(tmp_path, case):
    reference_or_skip()
    from nessai.livepoint import empty_structured_array

    from nessai_b200.nessai_plugin import diagonal_rescaling

    prop, model, live = make_proposal(tmp_path, **CASES[case])
    out = diagonal_rescaling(prop._reparameterisation, prop.prime_parameters, model.names)
    assert out is not None, case
    scale, shift = out
    rng = np.random.default_rng(0)
    xp = empty_structured_array(64, names=prop.prime_parameters)
    a = rng.uniform(-1.5, 1.5, size=(64, D))
    for i, p in enumerate(prop.prime_parameters):
        xp[p] = a[:, i]
    x_ref, log_j_ref = prop.inverse_rescale(xp.copy())
    got = a * scale + shift
    ref = np.stack([x_ref[n] for n in model.names], axis=-1)
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(log_j_ref, np.sum(np.log(np.abs(scale))), rtol=1e-13, atol=1e-13)
    # and the forward direction the training data take
    x_prime, log_j_fwd = prop.rescale(live.copy())
    back = np.stack([x_prime[p] for p in prop.prime_parameters], axis=-1) * scale + shift
    np.testing.assert_allclose(back, np.stack([live[n] for n in model.names], axis=-1), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(log_j_fwd, -np.sum(np.log(np.abs(scale))), rtol=1e-13, atol=1e-13)


REFUSED = {
    "logit": dict(reparameterisations={"logit": dict(parameters=[f"x{i}" for i in range(D)])}),
    "one_logit": dict(reparameterisations={"x0": "logit", "x1": "default", "x2": "default", "x3": "default"}),
}


@pytest.mark.parametrize("case", list(REFUSED))
def test_not_a_diagonal_affine_is_refused(tmp_path, case):
    reference_or_skip()
    from nessai_b200.nessai_plugin import diagonal_rescaling

    prop, model, _ = make_proposal(tmp_path, **REFUSED[case])
    assert diagonal_rescaling(prop._reparameterisation, prop.prime_parameters, model.names) is None
