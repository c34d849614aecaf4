"""Transform base classes (oracle only).

Serves /root/reference/src/nessai/flows/base.py:189-221 (NFlow delegates to
``CompositeTransform.forward/inverse``) and realnvp.py:212, nsf.py:128.
"""

import torch
from torch import nn


class InverseNotAvailable(Exception):
    """Raised when a transform has no (available) inverse."""


class InputOutsideDomain(Exception):
    """Raised when the input to a transform is outside its domain."""


class Transform(nn.Module):
    """Base class: ``forward``/``inverse`` return ``(outputs, logabsdet)``."""

    def forward(self, inputs, context=None):
        raise NotImplementedError()

    def inverse(self, inputs, context=None):
        raise InverseNotAvailable()


class CompositeTransform(Transform):
    """Apply transforms in order, summing log|det J| from ``zeros(N)``."""

    def __init__(self, transforms):
        super().__init__()
        self._transforms = nn.ModuleList(transforms)

    @staticmethod
    def _cascade(inputs, funcs, context):
        outputs = inputs
        total_logabsdet = inputs.new_zeros(inputs.shape[0])
        for func in funcs:
            outputs, logabsdet = func(outputs, context)
            total_logabsdet = total_logabsdet + logabsdet
        return outputs, total_logabsdet

    def forward(self, inputs, context=None):
        return self._cascade(inputs, self._transforms, context)

    def inverse(self, inputs, context=None):
        funcs = (t.inverse for t in self._transforms[::-1])
        return self._cascade(inputs, funcs, context)


class InverseTransform(Transform):
    """Swap forward and inverse of a transform."""

    def __init__(self, transform):
        super().__init__()
        self._transform = transform

    def forward(self, inputs, context=None):
        return self._transform.inverse(inputs, context)

    def inverse(self, inputs, context=None):
        return self._transform(inputs, context)


__all__ = [
    "CompositeTransform",
    "InputOutsideDomain",
    "InverseNotAvailable",
    "InverseTransform",
    "Transform",
    "torch",
]
