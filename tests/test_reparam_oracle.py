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
from conftest import REPO, reference_or_skip, simt_or_skip

D = 4
NAMES = [f"x{i}" for i in range(D)]


def make_proposal(tmp_path, box=None, **kw):
    from nessai.livepoint import numpy_array_to_live_points
    from nessai.model import Model
    from nessai.proposal import FlowProposal

    class Box(Model):
        def __init__(self):
            self.names = list(NAMES)
            self.bounds = ({n: list(box) for n in self.names} if box
                           else {n: [-4.0 - i, 6.0 + 2 * i] for i, n in enumerate(self.names)})

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
    pts = (rng.uniform(box[0] + 0.05 * (box[1] - box[0]), box[1] - 0.05 * (box[1] - box[0]), (300, D)) if box
           else 1.3 * rng.standard_normal((300, D)) + 0.4)
    live = numpy_array_to_live_points(pts, model.names)
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
    # ScaleAndShift with a post-rescaling: x = Q^-1(x') * scale + shift
    "zscore_gaussian_cdf": (dict(reparameterisations={"zscore-gaussian-cdf": dict(parameters=NAMES)}), None, {6}),
    # ... and with a pre-rescaling: x = P^-1(scale * x' + shift) (needs x in P's domain: see BOX)
    "zscore_logit": (dict(reparameterisations={"z-score-logit": dict(parameters=NAMES)}), None, {1}),
    "zscore_inv_gaussian_cdf": (dict(reparameterisations={"z-score-inv-gaussian-cdf": dict(parameters=NAMES)}),
                                None, {5}),
    "log_zscore": (dict(reparameterisations={"log-z-score": dict(parameters=NAMES)}), None, {3}),
    "rescale_pre_log_post_none": (dict(reparameterisations={"rescaletobounds": dict(
        parameters=NAMES, pre_rescaling="log", update_bounds=False)}), None, {3}),
    "rescale_post_exp": (dict(reparameterisations={"rescaletobounds": dict(
        parameters=NAMES, post_rescaling="exp", update_bounds=False, rescale_bounds=[1.0, 2.0])}), None, {4}),
    "rescale_post_inv_gaussian_cdf": (dict(reparameterisations={"rescaletobounds": dict(
        parameters=NAMES, post_rescaling="inv_gaussian_cdf", update_bounds=False, rescale_bounds=[0.0, 1.0])}),
        None, {5}),
}
# configurations whose forward map needs the physical parameters inside the function's domain
BOX = {"zscore_logit": (0.0, 1.0), "zscore_inv_gaussian_cdf": (0.0, 1.0), "log_zscore": (0.0, 1.0),
       "rescale_pre_log_post_none": (0.5, 3.0)}


@pytest.mark.reference
@pytest.mark.parametrize("case", list(CASES))
def test_parameter_maps_and_oracle_match_reference_inverse_rescale(tmp_path, case):
    reference_or_skip()
    from nessai.livepoint import empty_structured_array

    from nessai_b200.nessai_plugin import diagonal_rescaling, parameter_maps
    from oracle.reparam_numpy import inverse_maps

    kw, edges, kinds = CASES[case]
    prop, model, live = make_proposal(tmp_path, box=BOX.get(case), **kw)
    if edges is not None:
        (r,) = [r for r in prop._reparameterisation.values()]
        for p, e in zip(NAMES, edges):
            r._edges[p] = e  # every branch of rescale.py:570-590, whatever the data suggested
    maps = parameter_maps(prop._reparameterisation, prop.prime_parameters, model.names)
    assert maps is not None
    kind, scale, shift, pre_scale, pre_shift = maps[:5]
    assert set(kind.tolist()) == kinds and maps.has_pre_affine == (case in BOX)
    assert maps.names == list(model.names) and np.array_equal(maps.src, np.stack([np.arange(D)] * 3, axis=1))
    assert diagonal_rescaling(prop._reparameterisation, prop.prime_parameters, model.names) is None
    rng = np.random.default_rng(0)
    n = 200
    a = rng.normal(0.0, 1.2, size=(n, D))
    if case == "log":
        a = -np.abs(a)  # log of a value in [0, 1]
    elif case == "rescale_post_exp":
        a = np.exp(rng.uniform(0.0, 0.7, size=(n, D)))  # exp of a value in [1, 2]
    elif case in ("zscore_gaussian_cdf", "rescale_post_inv_gaussian_cdf"):
        a = rng.uniform(0.01, 0.99, size=(n, D)) if case == "zscore_gaussian_cdf" else a
    xp = empty_structured_array(n, names=prop.prime_parameters)
    for i, p in enumerate(prop.prime_parameters):
        xp[p] = a[:, i]
    x_ref, log_j_ref = prop.inverse_rescale(xp.copy())
    x, log_j = inverse_maps(a, kind, scale, shift, pre_scale, pre_shift)
    ref = np.stack([x_ref[nm] for nm in model.names], axis=-1)
    np.testing.assert_allclose(x, ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(log_j, log_j_ref, rtol=1e-12, atol=1e-12)
    # and the forward direction (what the training data take) is its inverse
    x_prime, log_j_fwd = prop.rescale(live.copy())
    back, lj_back = inverse_maps(np.stack([x_prime[p] for p in prop.prime_parameters], axis=-1), kind, scale, shift,
                                 pre_scale, pre_shift)
    if edges is None:  # (forced edges are not the ones rescale() used)
        np.testing.assert_allclose(back, np.stack([live[nm] for nm in model.names], axis=-1), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(lj_back, -log_j_fwd, rtol=1e-9, atol=1e-9)


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
    lib.tail_rows_host.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 9 + [C.c_double, C.c_double] + [C.c_void_p] * 4
    lib.nb200_host_erfcinv.restype = C.c_double
    lib.nb200_host_erfcinv.argtypes = [C.c_double]
    return lib


def test_host_erfcinv_helper(host_tail):
    """The harness's stand-in for CUDA's erfcinv is accurate to double precision."""
    from scipy.special import erfcinv

    y = np.concatenate([np.logspace(-300, -1, 400), np.linspace(0.1, 1.9, 400), 2.0 - np.logspace(-15, -1, 100)])
    got = np.array([host_tail.nb200_host_erfcinv(float(v)) for v in y])
    np.testing.assert_allclose(got, erfcinv(y), rtol=2e-14, atol=2e-15)
    assert host_tail.nb200_host_erfcinv(0.0) == np.inf and host_tail.nb200_host_erfcinv(2.0) == -np.inf
    assert np.isnan(host_tail.nb200_host_erfcinv(-0.5)) and np.isnan(host_tail.nb200_host_erfcinv(2.5))


PI = np.pi
TAIL_CASE = dict(
    # slots 0 .. 11: the single-feature kinds; 12 .. 15: Angle (angle, aux radius, radius, angle mod 2 pi);
    # 16: floor (Dequantise); 17, 18: ToCartesian + its auxiliary radius; 19 .. 21: AnglePair az-zen with an
    # auxiliary radius; 22 .. 24: AnglePair ra-dec with a radial parameter; 25: an augment parameter
    # (slot 1: floor of the final value, the 0x100 flag of "dequantise-logit")
    kind=np.array([0, 1 | 0x100, 2, 3, 1, 2, 0, 4, 5, 6, 1, 3, 7, 10, 9, 8, 11, 12, 10, 8, 13, 16, 7, 14, 15, 17],
                  dtype=np.int32),
    scale=np.array([1.5, 8.0, -3.0, 2.0, 0.5, 4.0, -0.7, 1.2, 6.0, 0.9, 1.0, 1.0, 0.5, 1.0, 1.0, 1.0,
                    1.0, 2.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]),
    shift=np.array([0.2, -4.0, 5.0, -1.0, 0.0, -2.0, 0.3, 0.1, -3.0, 0.4, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
                    0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]),
    lo=np.array([-3.0, -4.0, 2.0, -1.0, 0.0, -2.0, -2.0, -3.0, -3.0, -2.0, 0.0, 0.0, -1.5, -np.inf, 0.0, 0.0,
                 -6.0, -1.0, -np.inf, 0.0, 0.0, -np.inf, -PI, -PI / 2, 0.0, -np.inf]),
    hi=np.array([3.0, 4.0, 5.0, 9.0, 0.5, 1.5, 2.0, 2.0, 3.0, 2.5, 1.0, 9.0, 1.5, np.inf, 3.5, 2 * PI,
                 9.0, 1.0, np.inf, 2 * PI, PI, np.inf, PI, PI / 2, 4.0, np.inf]),
    pre_scale=np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.3, 1.0, 0.2, 1.7, 0.6, 1.0, 1.0, 1.0, 1.0,
                        3.0, 1.0 / PI, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]),
    pre_shift=np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.5, 0.0, 0.5, -0.3, 0.2, 0.0, 0.0, 0.0, 0.0,
                        2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]),
    # slots 12 / 13 read the flow features (13, 12) as (x', y'), slots 14 / 15 the features (14, 15);
    # slots 2 and 5 are swapped to exercise the permutation of single-feature kinds; the triples read
    # (19, 20, 21) and, permuted, (24, 22, 23)
    src=np.array([[0, 0, 0], [1, 1, 1], [5, 5, 5], [3, 3, 3], [4, 4, 4], [2, 2, 2], [6, 6, 6], [7, 7, 7], [8, 8, 8],
                  [9, 9, 9], [10, 10, 10], [11, 11, 11], [13, 12, 13], [13, 12, 13], [14, 15, 14], [14, 15, 14],
                  [16, 16, 16], [17, 18, 17], [17, 18, 17], [19, 20, 21], [19, 20, 21], [19, 20, 21],
                  [24, 22, 23], [24, 22, 23], [24, 22, 23], [25, 25, 25]], dtype=np.int32),
)


def tail_case_inputs(n, seed=5):
    """Rows for TAIL_CASE: every kind, saturated / overflowing / out-of-domain arguments and
    rows the draw kernel has already dropped."""
    rng = np.random.default_rng(seed)
    d = len(TAIL_CASE["kind"])
    xp = rng.normal(0.0, 1.0, size=(n, d)).astype(np.float32)
    if n >= 200:
        xp[:50, 1] = rng.choice([-60.0, 45.0, 800.0, -800.0], size=50)  # saturated sigmoid: log|J| = -inf
        xp[50:80, 3] = 900.0  # exp overflow
        xp[80:110, 7] = -6.0  # log of a negative number: NaN
        xp[110:140, 9] = rng.choice([-4.0, 4.0], size=30)  # quantile outside (0, 1): NaN
        xp[140:170, 8] = rng.choice([-40.0, 40.0], size=30)  # normal CDF saturates at 0 / 1
    logq_flow = rng.normal(-8.0, 2.0, size=n)
    logq_flow[rng.random(n) < 0.1] = np.nan  # rows the draw kernel already dropped
    return xp, logq_flow


@pytest.mark.parametrize("min_log_q,pre", [(None, True), (-14.0, True), (None, False)])
def test_kernel_row_function_matches_oracle(host_tail, min_log_q, pre):
    from oracle.reparam_numpy import tail_rows

    n = 5000
    c = {k: v.copy() for k, v in TAIL_CASE.items()}
    if not pre:  # ... and without the optional arrays: slot d reads feature d
        c["pre_scale"], c["pre_shift"], c["src"] = None, None, None
        c["kind"][[7, 9]] = 0  # their arguments rely on the pre-affine map to be in the domain
    d = len(c["kind"])
    xp, logq_flow = tail_case_inputs(n)
    x_ref, lq_ref, lw_ref, valid = tail_rows(xp, logq_flow, log_prior_const=-2.5, min_log_q=min_log_q, **c)
    logq, logw = logq_flow.copy(), np.empty(n)
    x64 = np.empty((n, d))
    stats = np.array([-np.inf, 0.0])
    ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    host_tail.tail_rows_host(n, d, xp.ctypes.data, c["kind"].ctypes.data, ptr(c["src"]), ptr(c["pre_scale"]),
                             ptr(c["pre_shift"]),
                             c["scale"].ctypes.data, c["shift"].ctypes.data, c["lo"].ctypes.data, c["hi"].ctypes.data,
                             -2.5, -np.inf if min_log_q is None else min_log_q,
                             logq.ctypes.data, logw.ctypes.data, x64.ctypes.data, stats.ctypes.data)
    assert 0.05 * n < valid.sum() < 0.9 * n
    np.testing.assert_array_equal(~np.isnan(logw), valid)
    np.testing.assert_array_equal(~np.isnan(logq), valid)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(x64, x_ref, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(logq[valid], lq_ref[valid], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(logw[valid], lw_ref[valid], rtol=1e-12, atol=1e-12)
    assert stats[1] == valid.sum() and stats[0] == logw[valid].max()


# ------------------------------------------------------------------ Angle (pair maps)
ANGLE_CASES = {
    # name: (model bounds of the angle, reparameterisations, expected kinds in x-space order)
    "angle_aux_radius": ((-np.pi, np.pi), {"a": "angle", "x": "default"}, [7, 0, 10]),
    "angle_zero_bound": ((0.0, 2 * np.pi), {"a": "angle", "x": "z-score"}, [8, 0, 10]),
    "angle_pi": ((0.0, np.pi), {"x": "logit", "a": "angle-pi"}, [8, 1, 10]),
    "periodic": ((0.0, 3.0), {"a": "periodic", "x": "default"}, [8, 0, 10]),  # scale = 2 pi / 3
    "angle_with_radial_parameter": ((0.0, 2 * np.pi), {"angle": {"parameters": ["a", "x"]}}, [8, 9]),
}


@pytest.mark.reference
@pytest.mark.parametrize("case", list(ANGLE_CASES))
def test_angle_maps_and_oracle_match_reference(tmp_path, case):
    """``Angle`` (reparameterisations/angle.py:17-186): the prime space holds Cartesian pairs; the
    x-space gains an auxiliary radius with a chi(2) prior unless the model has a radial parameter.
    ``parameter_maps`` + the oracle against the reference's ``inverse_rescale`` and the
    reparameterisation's own ``log_prior``."""
    reference_or_skip()
    from nessai.livepoint import empty_structured_array, numpy_array_to_live_points
    from nessai.model import Model
    from nessai.proposal import FlowProposal

    from nessai_b200.nessai_plugin import diagonal_rescaling, parameter_maps
    from oracle.reparam_numpy import inverse_maps

    bounds, reparams, kinds = ANGLE_CASES[case]

    class M(Model):
        def __init__(self):
            self.names = ["a", "x"]
            self.bounds = {"a": list(bounds), "x": [0.5, 4.0]}

        def log_prior(self, x):
            return np.log(self.in_bounds(x), dtype="float")

        def log_likelihood(self, x):
            return np.zeros(x.size)

    model = M()
    rng = np.random.default_rng(4)
    model.set_rng(rng)
    prop = FlowProposal(model, rng=rng, flow_config=dict(n_blocks=2, n_neurons=8), output=str(tmp_path), poolsize=100,
                        plot=False, reparameterisations=reparams)
    prop.initialise()
    live = numpy_array_to_live_points(np.stack([rng.uniform(*bounds, 300), rng.uniform(0.6, 3.9, 300)], axis=1),
                                      model.names)
    prop.check_state(live)
    maps = parameter_maps(prop._reparameterisation, prop.prime_parameters, model.names, prop.parameters)
    assert maps is not None and maps.names == list(prop.parameters) and not maps.affine
    assert maps.kind.tolist() == kinds
    assert diagonal_rescaling(prop._reparameterisation, prop.prime_parameters, model.names) is None
    n, Dp = 400, len(prop.prime_parameters)
    assert Dp == len(prop.parameters)
    a = rng.normal(0.0, 1.0, size=(n, Dp))
    xp = empty_structured_array(n, names=prop.prime_parameters)
    for i, p in enumerate(prop.prime_parameters):
        xp[p] = a[:, i]
    x_ref, log_j_ref = prop.inverse_rescale(xp.copy())
    x, log_j, log_p = inverse_maps(a, maps.kind, maps.scale, maps.shift, maps.pre_scale, maps.pre_shift, maps.src,
                                   return_log_prior=True)
    ref = np.stack([x_ref[nm] for nm in prop.parameters], axis=-1)
    np.testing.assert_allclose(x, ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(log_j, log_j_ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(log_p, prop._reparameterisation.log_prior(x_ref), rtol=1e-12, atol=1e-12)
    # forward (training data) then inverse gives the angle back
    x_prime, _ = prop.rescale(live.copy())
    back = inverse_maps(np.stack([x_prime[p] for p in prop.prime_parameters], axis=-1), maps.kind, maps.scale,
                        maps.shift, maps.pre_scale, maps.pre_shift, maps.src)[0]
    np.testing.assert_allclose(back[:, 0], live["a"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(back[:, 1], live["x"], rtol=1e-9, atol=1e-9)


PAIR_CASES = {
    # name: (model names, bounds, reparameterisations, expected kinds in x-space order)
    "to_cartesian": (["a", "x"], {"a": [0.5, 3.0], "x": [0.5, 4.0]}, {"a": "to-cartesian", "x": "default"},
                     [12, 0, 10]),
    "angle_pair_az_zen_aux_radius": (["az", "zen", "x"], {"az": [0.0, 2 * np.pi], "zen": [0.0, np.pi], "x": [0.5, 4.0]},
                                     {"angle-pair": {"parameters": ["az", "zen"]}, "x": "default"}, [8, 13, 0, 16]),
    "angle_pair_ra_dec_radial": (["d", "dec", "ra"],
                                 {"ra": [-np.pi, np.pi], "dec": [-np.pi / 2, np.pi / 2], "d": [0.5, 4.0]},
                                 {"angle-pair": {"parameters": ["ra", "dec", "d"]}}, [15, 14, 7]),
    "angle_pair_sky_zero_bound": (["ra", "dec"], {"ra": [0.0, 2 * np.pi], "dec": [-np.pi / 2, np.pi / 2]},
                                  {"angle-pair": {"parameters": ["ra", "dec"]}}, [8, 14, 16]),
    "dequantise": (["k", "x"], {"k": [0.0, 5.0], "x": [0.5, 4.0]}, {"k": "dequantise", "x": "default"}, [11, 0]),
    "dequantise_logit": (["k", "x"], {"k": [0.0, 5.0], "x": [0.5, 4.0]}, {"k": "dequantise-logit", "x": "default"},
                         [1 | 0x100, 0]),
}


@pytest.mark.reference
@pytest.mark.parametrize("case", list(PAIR_CASES))
def test_cartesian_pair_and_dequantise_maps_match_reference(tmp_path, case):
    """``ToCartesian`` (angle.py:189-232), ``AnglePair`` (angle.py:235-538, both conventions, with a
    radial parameter or the auxiliary chi(3) one) and ``Dequantise`` (discrete.py): ``parameter_maps``
    + the oracle against the reference's ``inverse_rescale`` and the reparameterisation's own
    ``log_prior``; the forward map of the reference, then our inverse, gives the parameters back."""
    reference_or_skip()
    from nessai.livepoint import empty_structured_array, numpy_array_to_live_points
    from nessai.model import Model
    from nessai.proposal import FlowProposal

    from nessai_b200.nessai_plugin import parameter_maps
    from oracle.reparam_numpy import inverse_maps

    names, bounds, reparams, kinds = PAIR_CASES[case]

    class M(Model):
        def __init__(self):
            self.names = list(names)
            self.bounds = {k: list(v) for k, v in bounds.items()}

        def log_prior(self, x):
            return np.log(self.in_bounds(x), dtype="float")

        def log_likelihood(self, x):
            return np.zeros(x.size)

        def new_point(self, N=1):
            x = super().new_point(N)
            if "k" in self.names:  # a discrete parameter (the proposal verifies its rescaling on these draws)
                x["k"] = np.floor(x["k"])
            return x

    model = M()
    rng = np.random.default_rng(4)
    model.set_rng(rng)
    prop = FlowProposal(model, rng=rng, flow_config=dict(n_blocks=2, n_neurons=8), output=str(tmp_path), poolsize=100,
                        plot=False, reparameterisations=reparams)
    prop.initialise()
    cols = []
    for nm in names:
        lo, hi = bounds[nm]
        v = rng.uniform(lo + 0.02 * (hi - lo), hi - 0.02 * (hi - lo), 300)
        cols.append(np.floor(rng.uniform(lo, hi + 1.0, 300)) if nm == "k" else v)
    live = numpy_array_to_live_points(np.stack(cols, axis=1), model.names)
    prop.check_state(live)
    maps = parameter_maps(prop._reparameterisation, prop.prime_parameters, model.names, prop.parameters)
    assert maps is not None and maps.names == list(prop.parameters) and not maps.affine
    assert maps.kind.tolist() == kinds
    n, Dp = 500, len(prop.prime_parameters)
    assert Dp == len(prop.parameters)
    a = rng.normal(0.0, 1.0, size=(n, Dp))
    xp = empty_structured_array(n, names=prop.prime_parameters)
    for i, p in enumerate(prop.prime_parameters):
        xp[p] = a[:, i]
    x_ref, log_j_ref = prop.inverse_rescale(xp.copy())
    x, log_j, log_p = inverse_maps(a, maps.kind, maps.scale, maps.shift, maps.pre_scale, maps.pre_shift, maps.src,
                                   return_log_prior=True)
    ref = np.stack([x_ref[nm] for nm in prop.parameters], axis=-1)
    np.testing.assert_allclose(x, ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(log_j, log_j_ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(log_p, prop._reparameterisation.log_prior(x_ref), rtol=1e-12, atol=1e-12)
    assert bool(log_p.any()) == any(r.has_prior for r in prop._reparameterisation.values())
    # forward (training data) then inverse gives the parameters back
    x_prime, _ = prop.rescale(live.copy())
    back = inverse_maps(np.stack([x_prime[p] for p in prop.prime_parameters], axis=-1), maps.kind, maps.scale,
                        maps.shift, maps.pre_scale, maps.pre_shift, maps.src)[0]
    for i, nm in enumerate(names):
        np.testing.assert_allclose(back[:, i], live[nm], rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------ the CUDA kernels under a CPU SIMT shim
@pytest.fixture(scope="module")
def simt_kernels(tmp_path_factory):
    """reparam_tail_kernel / sum_exp_kernel -- the CUDA source, unchanged -- compiled by g++ against
    tests/_hostcheck/simt_shim.h (one OS thread per CUDA thread, barriers for __syncthreads and the
    warp shuffles)."""
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("simt") / "libreparam_simt.so"
    d = os.path.join(REPO, "tests", "_hostcheck")
    res = subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-o", str(out),
                          os.path.join(d, "reparam_kernels_simt.cpp"), os.path.join(d, "reparam_host.cpp")],
                         capture_output=True, text=True)
    if res.returncode != 0:
        if "barrier" in res.stderr:
            pytest.skip("this g++ has no <barrier>")
        raise RuntimeError(res.stderr)
    lib = C.CDLL(str(out))
    lib.simt_reparam_tail.restype = None
    lib.simt_reparam_tail.argtypes = ([C.c_int, C.c_int64, C.c_int] + [C.c_void_p] * 9 + [C.c_double, C.c_double]
                                      + [C.c_void_p] * 4)
    lib.simt_sum_exp.restype = None
    lib.simt_sum_exp.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    return simt_or_skip(lib, 256)


@pytest.mark.parametrize("n,grid,pre", [(1500, 3, True), (700, 4, False), (5, 1, True), (256, 1, True)])
def test_tail_kernel_source_runs_under_simt_shim(simt_kernels, n, grid, pre):
    """The whole kernel, not only its row function: constants staged in shared memory, the
    grid-stride loop (n > grid * 256 and n < grid * 256), the warp reduction and the published
    statistics, against the oracle."""
    from oracle.reparam_numpy import tail_rows

    c = {k: v.copy() for k, v in TAIL_CASE.items()}
    if not pre:
        c["pre_scale"], c["pre_shift"], c["src"] = None, None, None
        c["kind"][[7, 9]] = 0
    d = len(c["kind"])
    xp, logq_flow = tail_case_inputs(n)
    x_ref, lq_ref, lw_ref, valid = tail_rows(xp, logq_flow, log_prior_const=-2.5, min_log_q=-14.0, **c)
    logq, logw = logq_flow.copy(), np.full(n, 123.0)
    x64 = np.full((n, d), 123.0)
    stats = np.array([-np.inf, 0.0])
    ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    simt_kernels.simt_reparam_tail(grid, n, d, xp.ctypes.data, c["kind"].ctypes.data, ptr(c["src"]),
                                   ptr(c["pre_scale"]), ptr(c["pre_shift"]), c["scale"].ctypes.data,
                                   c["shift"].ctypes.data, c["lo"].ctypes.data, c["hi"].ctypes.data, -2.5, -14.0,
                                   logq.ctypes.data, logw.ctypes.data, x64.ctypes.data, stats.ctypes.data)
    np.testing.assert_array_equal(~np.isnan(logw), valid)
    np.testing.assert_array_equal(~np.isnan(logq), valid)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(x64, x_ref, rtol=1e-13, atol=1e-13)  # every row written, none twice
    np.testing.assert_allclose(logq[valid], lq_ref[valid], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(logw[valid], lw_ref[valid], rtol=1e-12, atol=1e-12)
    assert stats[1] == valid.sum()
    if valid.any():
        assert stats[0] == logw[valid].max()
    else:
        assert stats[0] == -np.inf


@pytest.mark.parametrize("n,grid", [(3000, 5), (100, 2), (0, 3)])
def test_sum_exp_kernel_source_runs_under_simt_shim(simt_kernels, n, grid):
    rng = np.random.default_rng(8)
    logw = rng.normal(-3.0, 1.0, size=max(n, 1))
    logw[rng.random(len(logw)) < 0.3] = np.nan
    if n > 10:
        logw[3] = -np.inf
    valid = ~np.isnan(logw[:n]) & (logw[:n] > -np.inf)
    mx = np.array([logw[:n][valid].max() if valid.any() else -np.inf, 0.0])
    partials = np.full(grid, np.nan)
    simt_kernels.simt_sum_exp(grid, logw.ctypes.data, n, mx.ctypes.data, partials.ctypes.data)
    assert not np.isnan(partials).any()  # every block writes its partial, zeros included
    np.testing.assert_allclose(partials.sum(), np.exp(logw[:n][valid] - mx[0]).sum() if valid.any() else 0.0,
                               rtol=1e-13)
    again = np.full(grid, np.nan)
    simt_kernels.simt_sum_exp(grid, logw.ctypes.data, n, mx.ctypes.data, again.ctypes.data)
    np.testing.assert_array_equal(partials, again)  # no atomics: reproducible


# ------------------------------------------------------------------ every named reparameterisation
@pytest.mark.reference
def test_every_named_reparameterisation_maps_to_the_device_tail(tmp_path):
    """Every name the reference registers (reparameterisations/__init__.py:40-200) is recognised by
    ``parameter_maps`` -- none of them falls back to the reference's host loop -- and its inverse
    agrees with the reference's ``inverse_rescale`` on random prime points."""
    reference_or_skip()
    from nessai.livepoint import empty_structured_array, numpy_array_to_live_points
    from nessai.model import Model
    from nessai.proposal import FlowProposal
    from nessai.reparameterisations import default_reparameterisations

    from nessai_b200.nessai_plugin import parameter_maps
    from oracle.reparam_numpy import inverse_maps

    registered = sorted(k for k in default_reparameterisations.keys() if isinstance(k, str))
    assert len(registered) >= 30
    unit = {"z-score-logit", "zscore-logit", "z-score-inv-gaussian-cdf", "zscore-inv-gaussian-cdf"}  # domain (0, 1)
    for name in registered:
        pair = name == "angle-pair"
        if pair:
            bounds = {"a": [0.0, 2 * np.pi], "b": [0.0, np.pi]}
        elif name in unit:
            bounds = {"a": [0.0, 1.0], "b": [0.5, 3.0]}
        elif name == "angle-pi":
            bounds = {"a": [0.0, np.pi], "b": [0.5, 3.0]}
        elif name.startswith("dequantise"):
            bounds = {"a": [0.0, 5.0], "b": [0.5, 3.0]}
        else:
            bounds = {"a": [0.0, 3.0] if name == "periodic" else [0.5, 3.0], "b": [0.5, 3.0]}

        class M(Model):
            def __init__(self):
                self.names = ["a", "b"]
                self.bounds = {k: list(v) for k, v in bounds.items()}

            def log_prior(self, x):
                return np.log(self.in_bounds(x), dtype="float")

            def log_likelihood(self, x):
                return np.zeros(x.size)

            def new_point(self, N=1):
                x = super().new_point(N)
                if name.startswith("dequantise"):
                    x["a"] = np.floor(x["a"])
                return x

        model = M()
        rng = np.random.default_rng(1)
        model.set_rng(rng)
        kwargs = {"scale": 2.0} if name in ("scale", "rescale", "scaleandshift") else {}
        rep = ({name: {"parameters": ["a", "b"]}} if pair
               else {name: {"parameters": ["a"], **kwargs}, "default": {"parameters": ["b"]}})
        prop = FlowProposal(model, rng=rng, flow_config=dict(n_blocks=2, n_neurons=8), output=str(tmp_path / name),
                            poolsize=100, plot=False, reparameterisations=rep)
        prop.initialise()
        cols = []
        for nm in model.names:
            lo, hi = bounds[nm]
            cols.append(rng.uniform(lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo), 300))
        if name.startswith("dequantise"):
            cols[0] = np.floor(rng.uniform(0.0, 6.0, 300))
        live = numpy_array_to_live_points(np.stack(cols, axis=1), model.names)
        prop.check_state(live)
        prop.rescale(live.copy())  # (boundary inversion detects its edges here)
        maps = parameter_maps(prop._reparameterisation, prop.prime_parameters, model.names, prop.parameters)
        assert maps is not None, f"{name} falls back to the host loop"
        n = 200
        a = rng.normal(0.0, 0.7, size=(n, len(prop.prime_parameters)))
        xp = empty_structured_array(n, names=prop.prime_parameters)
        for i, p in enumerate(prop.prime_parameters):
            xp[p] = a[:, i]
        with np.errstate(all="ignore"):
            x_ref, log_j_ref = prop.inverse_rescale(xp.copy())
            x, log_j = inverse_maps(a, maps.kind, maps.scale, maps.shift, maps.pre_scale, maps.pre_shift, maps.src)
        ref = np.stack([x_ref[nm] for nm in prop.parameters], axis=-1)
        ok = np.isfinite(log_j_ref) & np.all(np.isfinite(ref), axis=1)
        assert ok.mean() > 0.3, name
        np.testing.assert_allclose(x[ok], ref[ok], rtol=1e-11, atol=1e-11, err_msg=name)
        np.testing.assert_allclose(log_j[ok], log_j_ref[ok], rtol=1e-11, atol=1e-11, err_msg=name)
