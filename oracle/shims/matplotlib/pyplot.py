"""Stub of matplotlib.pyplot (oracle only)."""
from _anything import Anything

rcParams = {"lines.linewidth": 1.0}


def subplots(*args, **kwargs):
    return Anything(), Anything()


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return Anything()
