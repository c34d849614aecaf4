"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Minimal stand-in for the third-party ``glasflow`` package (unpinned dependency
of the reference, /root/reference/pyproject.toml:25) which bundles a fork of
``nflows`` as ``glasflow.nflows``.  Neither package is installed in this image
and there is no network, so the arithmetic the reference delegates to nflows
is RESTATED here in plain PyTorch from the published nflows (v0.14) algorithm,
following the call sites listed in SURVEY.md section 8(c).

PARITY UNPINNED at the nflows boundary: the reference's tests hold no golden
vectors for any flow (only self-consistency properties, which
tests/test_oracle_* re-check against this restatement).
"""

__version__ = "0.0.0+b200-oracle-shim"
