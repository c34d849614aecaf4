"""The oracle against the golden vectors and the properties the reference pins.

Golden vectors = outputs of the UNMODIFIED reference (FlowModel / configure_model)
on the restated nflows shim, frozen by tests/golden/make_golden.py.  The reference
has no known-answer vectors of its own for the flow; what its tests
do pin are the self-consistency properties re-checked here.
"""

import numpy as np
import pytest
from conftest import reference_or_skip

from oracle.flow_numpy import NumpyFlow
from oracle.philox_numpy import philox4x32_10


def numpy_flow(cfg, sd):
    return NumpyFlow(
        sd,
        ftype={"nsf": "nsf", "maf": "maf"}.get(str(cfg.get("ftype")).lower(), "realnvp"),
        net=cfg.get("net", "resnet"),
        activation_name=cfg.get("activation", "relu"),
        volume_preserving=cfg.get("use_volume_preserving", False),
        num_bins=cfg.get("num_bins", 8),
        tail_bound=cfg.get("tail_bound", 5.0),
        hidden_features=cfg["n_neurons"],
    )


def test_numpy_oracle_matches_reference_golden(golden):
    name, g, cfg, sd = golden
    nf = numpy_flow(cfg, sd)
    z, lj = nf.forward(g["x"])
    x, ilj = nf.inverse(g["z"])
    # fp32 reference vs float64 restatement: fp32 rounding only
    tol = 2e-4 if "nsf" in name else 5e-5
    np.testing.assert_allclose(z, g["fwd_z"], atol=tol, rtol=1e-4)
    np.testing.assert_allclose(lj, g["fwd_logj"], atol=tol, rtol=1e-4)
    np.testing.assert_allclose(x, g["inv_x"], atol=tol, rtol=1e-4)
    np.testing.assert_allclose(ilj, g["inv_logj"], atol=tol, rtol=1e-4)
    np.testing.assert_allclose(nf.log_prob(g["x"]), g["fwd_logprob"], atol=tol, rtol=1e-4)
    np.testing.assert_allclose(nf.sample_and_log_prob(g["z"])[1], g["inv_logq"], atol=tol, rtol=1e-4)


def test_invertibility_and_logj_sign(golden):
    """/root/reference/tests/test_flows/test_included_flows.py:144-154."""
    name, g, cfg, sd = golden
    nf = numpy_flow(cfg, sd)
    z, lj = nf.forward(g["x"])
    x, ilj = nf.inverse(z)
    np.testing.assert_allclose(x, g["x"], atol=1e-8)
    np.testing.assert_allclose(lj, -ilj, atol=1e-8)


def test_sample_and_log_prob_consistency(golden):
    """test_included_flows.py:129-141: log_prob(x(z)) == base(z) - logJ_inv."""
    name, g, cfg, sd = golden
    nf = numpy_flow(cfg, sd)
    x, lq = nf.sample_and_log_prob(g["z"])
    np.testing.assert_allclose(nf.log_prob(x), lq, atol=1e-8)


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, out in kat:
        assert tuple(int(v) for v in philox4x32_10(*ctr, *key)) == out


@pytest.mark.reference
def test_shim_runs_reference_flow_properties():
    """The reference's own classes on the shim: forward_and_log_prob is bit-equal
    to forward + log_prob (test_included_flows.py:114-126)."""
    reference_or_skip()
    import torch
    from nessai.flows import configure_model

    for cfg in (
        dict(n_inputs=4, n_neurons=8, n_blocks=2, n_layers=2, ftype="realnvp"),
        dict(n_inputs=4, n_neurons=8, n_blocks=2, n_layers=2, ftype="nsf"),
    ):
        torch.manual_seed(1)
        m = configure_model(cfg)
        m.eval()
        x = torch.randn(100, 4)
        with torch.inference_mode():
            z, lp = m.forward_and_log_prob(x)
            z2, _ = m.forward(x)
            lp2 = m.log_prob(x)
        assert torch.equal(z, z2) and torch.equal(lp, lp2)
