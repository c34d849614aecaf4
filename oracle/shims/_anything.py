"""Permissive stand-in object for absent plotting packages (oracle only).

matplotlib / seaborn / cycler are not installed in this image; the reference
imports them at module import time (/root/reference/src/nessai/plot.py:10-15)
but never calls them when ``plot=False``.  Any attribute access or call on an
``Anything`` returns another ``Anything``; it also works as a context manager,
a decorator and an iterable of nothing.
"""


class Anything:
    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Anything()

    def __call__(self, *args, **kwargs):
        return Anything()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __iter__(self):
        return iter(())

    def __add__(self, other):
        return self

    __radd__ = __iadd__ = __add__

    def __getitem__(self, key):
        return Anything()

    def __setitem__(self, key, value):
        pass
