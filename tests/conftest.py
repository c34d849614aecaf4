import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")
GOLDEN_NAMES = [
    "c2_realnvp_mlp",
    "c2_realnvp_resnet",
    "d5_realnvp_perm_tanh",
    "d4_realnvp_additive_silu",
    "c1_realnvp_2d",
    "d6_nsf",
    "d8_maf",
    "d5_realnvp_mvn",
    "c3_nsf_trained",
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "reference: needs baseline/_ref (the installed reference)")


def load_golden(name):
    import json

    import numpy as np

    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cfg = json.loads(str(g["flow_config"]))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    return g, cfg, sd


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return (request.param,) + load_golden(request.param)


def reference_or_skip():
    import oracle.refenv as refenv

    if not refenv.reference_available():
        pytest.skip("reference not installed under baseline/_ref")
    refenv.activate()


def simt_or_skip(lib, block=512):
    """A harness built on tests/_hostcheck/simt_shim.h runs one OS thread per CUDA thread: skip
    where the machine cannot hold a whole block of them."""
    if lib.simt_probe(int(block)) != 0:
        pytest.skip(f"cannot run {block} threads at once here (SIMT shim)")
    return lib
