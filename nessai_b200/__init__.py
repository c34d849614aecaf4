"""nessai_b200 -- B200-native flow-proposal hot path for nessai.

Host side mirrors the reference's plugin surface
(``FlowModel`` /root/reference/src/nessai/flowmodel/base.py:25,
``FlowProposal`` /root/reference/src/nessai/proposal/flowproposal/flowproposal.py:29)
above a C-ABI CUDA library (``include/nessai_b200.h``).  Importing this package
never creates a CUDA context (the reference forks its likelihood pool before any
flow exists, /root/reference/src/nessai/flowsampler.py:155-157).
"""

__version__ = "0.1.0"

from .spec import FlowSpec  # noqa: F401


def __getattr__(name):
    if name in ("B200FlowModel", "FlowModel"):
        from .flowmodel import B200FlowModel

        return B200FlowModel
    if name in ("B200FlowProposal", "FlowProposal"):
        from .proposal import B200FlowProposal

        return B200FlowProposal
    raise AttributeError(name)
