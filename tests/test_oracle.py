"""The oracle against the golden vectors and the properties the reference pins.

Golden vectors = outputs of the UNMODIFIED reference (FlowModel / configure_model)
on the restated nflows shim, frozen by tests/golden/make_golden.py.  The reference
has no known-answer vectors of its own for the flow; what its tests
do pin are the self-consistency properties re-checked here.
"""

import os

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
        base_var=(cfg.get("distribution_kwargs") or {}).get("var", 1.0),
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


# ------------------------------------------------------------------ the CUDA Philox source on the host
@pytest.fixture(scope="module")
def philox_source():
    """nessai_b200/csrc/philox.cuh -- the header the draw kernels include, its generator being
    ``__host__ __device__`` -- compiled for the host by g++ (tests/_hostcheck/philox_host.cpp)."""
    import ctypes as C
    import shutil
    import subprocess
    import tempfile

    from conftest import REPO

    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if gxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    out = os.path.join(tempfile.mkdtemp(), "libphilox_host.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{cuda_inc}", "-o", out,
                    os.path.join(REPO, "tests", "_hostcheck", "philox_host.cpp"), "-lm"], check=True)
    lib = C.CDLL(out)
    lib.philox_host.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.latent_row_host.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    lib.accept_uniform_host.argtypes = [C.c_uint64, C.c_uint64]
    lib.accept_uniform_host.restype = C.c_double
    return lib


def test_cuda_philox_source_known_answers(philox_source):
    """Random123 known-answer vectors through the CUDA header itself: counter = (row_lo, row_hi,
    block, stream), key = (seed_lo, seed_hi)."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    out = np.zeros(4, dtype=np.uint32)
    for ctr, key, expect in kat:
        philox_source.philox_host(key[0] | (key[1] << 32), ctr[0] | (ctr[1] << 32), ctr[2], ctr[3], out.ctypes.data)
        assert tuple(int(v) for v in out) == expect


def test_cuda_draw_source_matches_numpy_restatement(philox_source):
    """The latent draw (Box-Muller on the Philox words, fp32) and the rejection step's uniform as
    the kernels form them, against oracle/philox_numpy.py -- the restatement every GPU parity
    test is driven by."""
    from oracle.philox_numpy import accept_uniform, latent_normals

    seed = 0x1234_5678_9ABC_DEF1
    rows = np.array([0, 1, 2, 1000, 2**32 - 1, 2**32, 2**40 + 12345], dtype=np.uint64)
    for D in (2, 5, 16, 32):
        z = np.zeros(D, dtype=np.float32)
        ref = latent_normals(seed, rows, D)
        for i, r in enumerate(rows):
            philox_source.latent_row_host(seed, int(r), D, z.ctypes.data)
            np.testing.assert_allclose(z, ref[i], rtol=2e-6, atol=2e-6)  # fp32 evaluation vs float64
    u = accept_uniform(seed, rows)
    got = [philox_source.accept_uniform_host(seed, int(r)) for r in rows]
    np.testing.assert_array_equal(got, u)
    assert np.all((u > 0) & (u < 1))


@pytest.mark.reference
@pytest.mark.parametrize("seed", range(12))
def test_numpy_oracle_matches_the_reference_over_the_config_space(seed):
    """The checker of the GPU sweeps (tests/test_gpu_fuzz.py) is itself pinned here: the float64 numpy
    oracle against the UNMODIFIED reference's own flow classes (configure_model on the glasflow shim,
    fp32) for random configurations over the same space -- incl. 17 .. 32-feature RealNVP flows, the
    default conditioner width, three residual blocks, tanh / SiLU -- with perturbed weights."""
    reference_or_skip()
    import torch
    from nessai.flows import configure_model
    from nessai.flowmodel.utils import update_flow_config
    from test_gpu_fuzz import draw_config

    cfg = draw_config(300 + seed)
    if seed % 3 == 0:  # make sure the 17 .. 32-feature RealNVP flows are in
        cfg = dict(n_inputs=int(17 + seed), n_blocks=3, n_layers=int(1 + seed % 3), ftype="realnvp")
    torch.manual_seed(seed)
    full = update_flow_config(dict(cfg))
    m = configure_model(dict(full))
    rng = np.random.default_rng(seed)
    sd = {}
    for k, v in m.state_dict().items():
        a = v.numpy().copy()
        if a.dtype.kind == "f" and not k.endswith(".mask"):
            a = a + (0.04 * rng.standard_normal(a.shape)).astype(np.float32)
            if "running_var" in k:
                a = np.abs(a) + 0.5
        sd[k] = a
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.eval()
    D = cfg["n_inputs"]
    z = rng.normal(size=(400, D)).astype(np.float32)
    with torch.inference_mode():
        x_ref, ilj_ref = m.inverse(torch.from_numpy(z))
        lp_ref = m.log_prob(x_ref)
    ocfg = dict(full)
    nf = numpy_flow(ocfg, sd)
    x, ilj = nf.inverse(z.astype(np.float64))
    ok = np.isfinite(ilj) & (np.abs(x).max(axis=1) < 50.0)
    assert ok.mean() > 0.9, cfg
    tol = dict(rtol=3e-4, atol=3e-4) if cfg["ftype"] in ("nsf", "maf") else dict(rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(x_ref.numpy()[ok], x[ok], err_msg=str(cfg), **tol)
    np.testing.assert_allclose(ilj_ref.numpy()[ok], ilj[ok], err_msg=str(cfg), **tol)
    np.testing.assert_allclose(lp_ref.numpy()[ok], nf.log_prob(x_ref.numpy().astype(np.float64))[ok], err_msg=str(cfg), **tol)
