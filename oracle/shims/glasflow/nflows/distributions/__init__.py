"""nflows.distributions restatement (oracle only)."""
from . import normal, uniform  # noqa: F401
from .base import Distribution, NoMeanException  # noqa: F401
from .normal import StandardNormal  # noqa: F401
from .uniform import BoxUniform  # noqa: F401
