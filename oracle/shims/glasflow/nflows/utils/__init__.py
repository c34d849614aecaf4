"""nflows.utils restatement (oracle only)."""
import torch

from . import torchutils  # noqa: F401
from .torchutils import sum_except_batch, tile, searchsorted  # noqa: F401


def create_alternating_binary_mask(features, even=True):
    """ones at ``start::2`` (start=0 if even else 1); ones = transformed.

    Call site: /root/reference/src/nessai/flows/nsf.py:99-101.
    """
    mask = torch.zeros(features).byte()
    start = 0 if even else 1
    mask[start::2] += 1
    return mask
