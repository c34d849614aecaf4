"""Stub of matplotlib (oracle only; see ../_anything.py)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(__file__)))
from _anything import Anything  # noqa: E402

from . import lines, pyplot  # noqa: E402,F401

__version__ = "0.0.0-stub"
rcParams = {}


def use(*args, **kwargs):
    pass


def rc_context(*args, **kwargs):
    return Anything()


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return Anything()
