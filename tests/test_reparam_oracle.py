"""The non-affine populate tail (logit / log post-rescaling, boundary inversion):

* ``oracle/reparam_numpy.py`` and ``nessai_plugin.parameter_maps`` pinned against the
  reference's own ``FlowProposal.inverse_rescale`` / ``rescale`` on the CPU;
* the per-row function of ``csrc/reparam_tail.cuh`` -- the same source the CUDA kernel runs --
  compiled for the host with g++ (tests/_hostcheck) and checked against the oracle."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
from conftest import REPO, reference_or_skip

D = 4
NAMES = [f"x{i}" for i in range(D)]


def make_proposal(tmp_path, **kw):
    from nessai.livepoint import numpy_array_to_live_points
    from nessai.model import Model
    from nessai.proposal import FlowProposal

    class Box(Model):
        def __init__(self):
            self.names = list(NAMES)
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
    prop.check_state(live)
    prop.rescale(live.copy())  # what train() does: boundary inversion detects its edges here
    return prop, model, live


CASES = {
    "logit": (dict(reparameterisations={"logit": dict(parameters=NAMES)}), None, {1}),
    "log": (dict(reparameterisations={"rescaletobounds": dict(parameters=NAMES, post_rescaling="log",
                                                               update_bounds=False)}), None, {3}),
    "one_logit_rest_mixed": (dict(reparameterisations={"x0": "logit", "x1": "default", "x2": "z-score",
                                                       "x3": "null"}), None, {0, 1}),
    "inversion_lower_upper": (dict(reparameterisations={"inversion": dict(parameters=NAMES)}),
                              ["lower", "upper", False, "lower"], {0, 2}),
    "inversion_offset": (dict(reparameterisations={"inversion": dict(parameters=NAMES, offset=True)}),
                         ["upper", "upper", "lower", False], {0, 2}),
}


@pytest.mark.reference
@pytest.mark.parametrize("case", list(CASES))
def test_parameter_maps_and_oracle_match_reference_inverse_rescale(tmp_path, case):
    reference_or_skip()
    from nessai.livepoint import empty_structured_array

    from nessai_b200.nessai_plugin import diagonal_rescaling, parameter_maps
    from oracle.reparam_numpy import inverse_maps

    kw, edges, kinds = CASES[case]
    prop, model, live = make_proposal(tmp_path, **kw)
    if edges is not None:
        (r,) = [r for r in prop._reparameterisation.values()]
        for p, e in zip(NAMES, edges):
            r._edges[p] = e  # every branch of rescale.py:570-590, whatever the data suggested
    maps = parameter_maps(prop._reparameterisation, prop.prime_parameters, model.names)
    assert maps is not None
    kind, scale, shift = maps
    assert set(kind.tolist()) == kinds
    assert diagonal_rescaling(prop._reparameterisation, prop.prime_parameters, model.names) is None
    rng = np.random.default_rng(0)
    n = 200
    a = rng.normal(0.0, 1.2, size=(n, D))
    if case == "log":
        a = -np.abs(a)  # log of a value in [0, 1]
    xp = empty_structured_array(n, names=prop.prime_parameters)
    for i, p in enumerate(prop.prime_parameters):
        xp[p] = a[:, i]
    x_ref, log_j_ref = prop.inverse_rescale(xp.copy())
    x, log_j = inverse_maps(a, kind, scale, shift)
    ref = np.stack([x_ref[nm] for nm in model.names], axis=-1)
    np.testing.assert_allclose(x, ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(log_j, log_j_ref, rtol=1e-12, atol=1e-12)


def test_undetected_edge_and_user_functions_are_refused(tmp_path):
    reference_or_skip()
    from nessai.utils.rescaling import logit, sigmoid

    from nessai_b200.nessai_plugin import parameter_maps

    prop, model, _ = make_proposal(tmp_path, reparameterisations={"inversion": dict(parameters=NAMES)})
    (r,) = prop._reparameterisation.values()
    r._edges[NAMES[0]] = None  # not detected yet (reset_inversion, rescale.py:662-665)
    assert parameter_maps(prop._reparameterisation, prop.prime_parameters, model.names) is None
    user = (lambda x: logit(x), lambda x: sigmoid(x))
    prop, model, _ = make_proposal(tmp_path, reparameterisations={
        "rescaletobounds": dict(parameters=NAMES, post_rescaling=user, update_bounds=False,
                                rescale_bounds=[0.0, 1.0])})
    assert parameter_maps(prop._reparameterisation, prop.prime_parameters, model.names) is None


# ------------------------------------------------------------------ the kernel's row function
@pytest.fixture(scope="module")
def host_tail(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("hostcheck") / "libreparam_host.so"
    src = os.path.join(REPO, "tests", "_hostcheck", "reparam_host.cpp")
    subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), src], check=True)
    lib = C.CDLL(str(out))
    lib.tail_rows_host.restype = None
    lib.tail_rows_host.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 6 + [C.c_double, C.c_double] + [C.c_void_p] * 4
    return lib


@pytest.mark.parametrize("min_log_q", [None, -12.0])
def test_kernel_row_function_matches_oracle(host_tail, min_log_q):
    from oracle.reparam_numpy import tail_rows

    rng = np.random.default_rng(5)
    n, d = 5000, 7
    kind = np.array([0, 1, 2, 3, 1, 2, 0], dtype=np.int32)
    scale = np.array([1.5, 8.0, -3.0, 2.0, 0.5, 4.0, -0.7])
    shift = np.array([0.2, -4.0, 5.0, -1.0, 0.0, -2.0, 0.3])
    lo = np.array([-3.0, -4.0, 2.0, -1.0, 0.0, -2.0, -2.0])
    hi = np.array([3.0, 4.0, 5.0, 9.0, 0.5, 1.5, 2.0])
    xp = rng.normal(0.0, 1.0, size=(n, d)).astype(np.float32)
    xp[:50, 1] = rng.choice([-60.0, 45.0, 800.0, -800.0], size=50)  # saturated sigmoid: log|J| = -inf
    xp[50:80, 3] = 900.0  # exp overflow
    logq_flow = rng.normal(-8.0, 2.0, size=n)
    logq_flow[rng.random(n) < 0.1] = np.nan  # rows the draw kernel already dropped
    x_ref, lq_ref, lw_ref, valid = tail_rows(xp, logq_flow, kind=kind, scale=scale, shift=shift, lo=lo, hi=hi,
                                             log_prior_const=-2.5, min_log_q=min_log_q)
    logq, logw = logq_flow.copy(), np.empty(n)
    x64 = np.empty((n, d))
    stats = np.array([-np.inf, 0.0])
    host_tail.tail_rows_host(n, d, xp.ctypes.data, kind.ctypes.data, scale.ctypes.data, shift.ctypes.data,
                             lo.ctypes.data, hi.ctypes.data, -2.5, -np.inf if min_log_q is None else min_log_q,
                             logq.ctypes.data, logw.ctypes.data, x64.ctypes.data, stats.ctypes.data)
    assert 0.05 * n < valid.sum() < 0.9 * n
    np.testing.assert_array_equal(~np.isnan(logw), valid)
    np.testing.assert_array_equal(~np.isnan(logq), valid)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(x64, x_ref, rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(logq[valid], lq_ref[valid], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(logw[valid], lw_ref[valid], rtol=1e-13, atol=1e-13)
    assert stats[1] == valid.sum() and stats[0] == logw[valid].max()
