"""nflows.transforms restatement (ORACLE ONLY -- see glasflow/__init__.py).

Every class cites the reference call site it serves (file:line under
/root/reference) and restates the published nflows v0.14 algorithm; the exact
formulas are the ones SURVEY.md section 8(c) lists.
"""

from . import normalization  # noqa: F401
from .autoregressive import MaskedAffineAutoregressiveTransform  # noqa: F401
from .base import (  # noqa: F401
    CompositeTransform,
    InputOutsideDomain,
    InverseNotAvailable,
    InverseTransform,
    Transform,
)
from .coupling import (  # noqa: F401
    AdditiveCouplingTransform,
    AffineCouplingTransform,
    CouplingTransform,
    PiecewiseRationalQuadraticCouplingTransform,
)
from .linear import Linear, LULinear, SVDLinear  # noqa: F401
from .nonlinearities import Logit, Sigmoid  # noqa: F401
from .normalization import ActNorm, BatchNorm  # noqa: F401
from .permutations import (  # noqa: F401
    Permutation,
    RandomPermutation,
    ReversePermutation,
)
