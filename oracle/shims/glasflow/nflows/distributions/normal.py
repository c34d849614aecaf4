"""StandardNormal (oracle only). Serves realnvp.py:208-209, nsf.py:124-125."""

import numpy as np
import torch

from ..utils import torchutils
from .base import Distribution


class StandardNormal(Distribution):
    """log p = -0.5 sum z^2 - 0.5 D log(2 pi); ``_log_z`` is a float64 buffer."""

    def __init__(self, shape):
        super().__init__()
        self._shape = torch.Size(shape)
        self.register_buffer(
            "_log_z",
            torch.tensor(
                0.5 * np.prod(shape) * np.log(2 * np.pi), dtype=torch.float64
            ),
            persistent=False,
        )

    def _log_prob(self, inputs, context):
        if inputs.shape[1:] != self._shape:
            raise ValueError(
                "Expected input of shape {}, got {}".format(
                    self._shape, inputs.shape[1:]
                )
            )
        neg_energy = -0.5 * torchutils.sum_except_batch(inputs**2, num_batch_dims=1)
        return neg_energy - self._log_z

    def _sample(self, num_samples, context):
        if context is None:
            return torch.randn(num_samples, *self._shape, device=self._log_z.device)
        raise NotImplementedError("oracle shim: no context support here")

    def _mean(self, context):
        if context is None:
            return self._log_z.new_zeros(self._shape)
        raise NotImplementedError()
