"""Make the UNMODIFIED reference importable (oracle / test infrastructure only).

``baseline/_ref`` holds ``nessai`` installed with
``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref
/root/reference`` (git-ignored, travels to the GPU box).  Its third-party
dependency ``glasflow`` and the plotting stack are absent from this image, so
``oracle/shims`` is appended for whichever of them cannot be imported.
"""

import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
REF_DIR = os.path.join(REPO, "baseline", "_ref")
SHIM_DIR = os.path.join(_HERE, "shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "nessai"))


def activate() -> None:
    """Idempotently put the reference + needed shims on ``sys.path``."""
    if not reference_available():
        raise RuntimeError(
            f"reference not installed under {REF_DIR}; run "
            "`python -c 'import __graft_entry__ as g; g.install_reference()'`"
        )
    needs_shim = any(
        importlib.util.find_spec(m) is None
        for m in ("glasflow", "matplotlib", "seaborn", "cycler")
    )
    if needs_shim and SHIM_DIR not in sys.path:
        # appended: a real install of any of these always wins
        sys.path.append(SHIM_DIR)
    if REF_DIR not in sys.path:
        sys.path.append(REF_DIR)


def using_shim() -> bool:
    import glasflow

    return "oracle-shim" in getattr(glasflow, "__version__", "")
