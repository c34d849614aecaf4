"""nflows.nn restatement (oracle only)."""
from . import nets  # noqa: F401
