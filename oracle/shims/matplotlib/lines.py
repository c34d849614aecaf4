"""Stub of matplotlib.lines (oracle only)."""
from _anything import Anything


class Line2D(Anything):
    pass
