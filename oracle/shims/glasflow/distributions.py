"""glasflow.distributions stand-ins (oracle only).

Serves /root/reference/src/nessai/flows/utils.py:14,51-57 and
flows/distributions.py:10-12,76-93.  Neither is used by any BASELINE config.
"""

import torch

from .nflows.distributions import Distribution


class MultivariateUniform(Distribution):
    """Uniform box distribution on [low, high]."""

    def __init__(self, low, high):
        super().__init__()
        low = torch.as_tensor(low)
        high = torch.as_tensor(high)
        if low.shape != high.shape:
            raise ValueError("low and high are not the same shape")
        if not (low < high).all():
            raise ValueError("low has elements that are higher than high")
        self._shape = low.shape
        self.register_buffer("low", low)
        self.register_buffer("high", high)
        self.register_buffer("_log_prob_value", -torch.sum(torch.log(high - low)))

    def _log_prob(self, inputs, context):
        lp = self._log_prob_value * inputs.new_ones(inputs.shape[0])
        inside = ((inputs >= self.low) & (inputs <= self.high)).all(dim=-1)
        return torch.where(inside, lp, lp.new_full(lp.shape, -float("inf")))

    def _sample(self, num_samples, context):
        u = torch.rand(num_samples, *self._shape, device=self.low.device)
        return self.low + u * (self.high - self.low)


class ResampledGaussian(Distribution):
    """LARS base distribution -- NOT restated."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "oracle shim: ResampledGaussian (LARS) is not restated"
        )

    def estimate_normalisation_constant(self, n_samples=1000, n_batches=1):
        raise NotImplementedError()
