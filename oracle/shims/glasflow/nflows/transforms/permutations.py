"""Permutation transforms (oracle only).

Serves /root/reference/src/nessai/flows/utils.py:290-291,307-311 and
maf.py:14-17,77-84.  ``_permutation`` is a buffer (lives in the state_dict).
"""

import torch

from .base import Transform


class Permutation(Transform):
    """Permute features along ``dim``; log|det J| = 0."""

    def __init__(self, permutation, dim=1):
        if permutation.ndimension() != 1:
            raise ValueError("Permutation must be a 1D tensor.")
        if not isinstance(dim, int) or dim < 1:
            raise ValueError("dim must be a positive integer.")
        super().__init__()
        self._dim = dim
        self.register_buffer("_permutation", permutation)

    @property
    def _inverse_permutation(self):
        return torch.argsort(self._permutation)

    @staticmethod
    def _permute(inputs, permutation, dim):
        if dim >= inputs.ndimension():
            raise ValueError("No dimension {} in inputs.".format(dim))
        if inputs.shape[dim] != len(permutation):
            raise ValueError(
                "Dimension {} in inputs must be of size {}.".format(
                    dim, len(permutation)
                )
            )
        outputs = torch.index_select(inputs, dim, permutation)
        logabsdet = inputs.new_zeros(inputs.shape[0])
        return outputs, logabsdet

    def forward(self, inputs, context=None):
        return self._permute(inputs, self._permutation, self._dim)

    def inverse(self, inputs, context=None):
        return self._permute(inputs, self._inverse_permutation, self._dim)


class RandomPermutation(Permutation):
    def __init__(self, features, dim=1):
        if not isinstance(features, int) or features < 1:
            raise ValueError("Number of features must be a positive integer.")
        super().__init__(torch.randperm(features), dim)


class ReversePermutation(Permutation):
    def __init__(self, features, dim=1):
        if not isinstance(features, int) or features < 1:
            raise ValueError("Number of features must be a positive integer.")
        super().__init__(torch.arange(features - 1, -1, -1), dim)
