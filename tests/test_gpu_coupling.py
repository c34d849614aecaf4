"""The stand-alone affine coupling transform against its numpy restatement
(oracle/flow_numpy.py::_coupling arithmetic with the conditioner output supplied)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def numpy_coupling(x, params, tf, additive, inverse):
    x = x.astype(np.float64)
    p = params.astype(np.float64)
    d_tr = len(tf)
    shift = p[:, :d_tr]
    scale = np.ones_like(shift) if additive else 1.0 / (1.0 + np.exp(-(p[:, d_tr:] + 2.0))) + 1e-3
    y = x.copy()
    if inverse:
        y[:, tf] = (x[:, tf] - shift) / scale
        ld = -np.log(scale).sum(1)
    else:
        y[:, tf] = x[:, tf] * scale + shift
        ld = np.log(scale).sum(1)
    return y, ld


@pytest.mark.parametrize("D,tf", [(16, list(range(1, 16, 2))), (16, list(range(0, 16, 2))), (16, [3, 0, 9]),
                                  (5, [1, 3]), (32, list(range(0, 32, 2))), (2, [1])])
@pytest.mark.parametrize("additive", [False, True])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("n", [1, 1001])
def test_coupling_transform_matches_numpy(D, tf, additive, inverse, n):
    from nessai_b200.coupling import coupling_transform

    rng = np.random.default_rng(D * 7 + n)
    x = rng.normal(size=(n, D)).astype(np.float32)
    params = rng.normal(size=(n, len(tf) * (1 if additive else 2))).astype(np.float32) * 2
    y, ld = coupling_transform(torch.from_numpy(x).cuda(), torch.from_numpy(params).cuda(), tf,
                               additive=additive, inverse=inverse)
    y64, ld64 = numpy_coupling(x, params, tf, additive, inverse)
    np.testing.assert_allclose(y.cpu().numpy(), y64, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(ld.cpu().numpy(), ld64, rtol=2e-5, atol=2e-5)


def test_coupling_round_trip_at_bench_size():
    """4e6 rows (the bench size): inverse(forward(x)) == x, logdets cancel."""
    from nessai_b200.coupling import coupling_transform

    n, D = 4_000_000, 16
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, D, device="cuda", generator=g)
    p = torch.randn(n, D, device="cuda", generator=g)
    tf = list(range(1, 16, 2))
    y, ld = coupling_transform(x, p, tf)
    xr, ldi = coupling_transform(y, p, tf, inverse=True)
    assert float((xr - x).abs().max()) < 1e-4
    assert float((ld + ldi).abs().max()) < 1e-5
    assert torch.equal(y[:, 0::2], x[:, 0::2])
