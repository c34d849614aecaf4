"""nflows.nn.nets restatement (oracle only)."""
from .resnet import ResidualBlock, ResidualNet  # noqa: F401
