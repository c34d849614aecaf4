"""The C-ABI library loads and exports every symbol include/nessai_b200.h declares
(no compute calls: there is no GPU on the CPU test tier)."""

import ctypes
import os
import re

from conftest import REPO


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from nessai_b200 import _lib

    header = open(os.path.join(REPO, "include", "nessai_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(nb200_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared)
    assert _lib.load().nb200_version() >= 100


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "nessai_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_missing_cuda_fails_loudly():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nessai_b200.flowmodel import B200FlowModel

    fm = B200FlowModel(dict(n_inputs=4, ftype="realnvp"), output="/tmp/nb200_t")
    with pytest.raises(RuntimeError):
        fm.initialise()
