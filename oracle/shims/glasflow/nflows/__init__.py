"""Restated subset of ``nflows`` (oracle only). See ``glasflow/__init__.py``."""
from . import distributions, nn, transforms, utils  # noqa: F401
