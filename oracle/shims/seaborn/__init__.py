"""Stub of seaborn (oracle only; see ../_anything.py)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(__file__)))
from _anything import Anything  # noqa: E402


def plotting_context(*args, **kwargs):
    return {}


def axes_style(*args, **kwargs):
    return Anything()


def color_palette(*args, **kwargs):
    return ["C0", "C1", "C2", "C3", "C4", "C5", "C6", "C7", "C8", "C9"]


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return Anything()
