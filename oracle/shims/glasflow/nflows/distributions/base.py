"""Distribution base class (oracle only). Serves flows/base.py:189,223-287."""

import torch
from torch import nn


class NoMeanException(Exception):
    pass


class Distribution(nn.Module):
    def forward(self, *args):
        raise RuntimeError("Forward method cannot be called for a Distribution object.")

    def log_prob(self, inputs, context=None):
        inputs = torch.as_tensor(inputs)
        if context is not None:
            context = torch.as_tensor(context)
            if inputs.shape[0] != context.shape[0]:
                raise ValueError(
                    "Number of input items must be equal to number of context items."
                )
        return self._log_prob(inputs, context)

    def _log_prob(self, inputs, context):
        raise NotImplementedError()

    def sample(self, num_samples, context=None, batch_size=None):
        if not isinstance(num_samples, int) or num_samples < 1:
            raise TypeError("Number of samples must be a positive integer.")
        if context is not None:
            context = torch.as_tensor(context)
        if batch_size is None:
            return self._sample(num_samples, context)
        num_batches = num_samples // batch_size
        num_leftover = num_samples % batch_size
        samples = [self._sample(batch_size, context) for _ in range(num_batches)]
        if num_leftover > 0:
            samples.append(self._sample(num_leftover, context))
        return torch.cat(samples, dim=0)

    def _sample(self, num_samples, context):
        raise NotImplementedError()

    def sample_and_log_prob(self, num_samples, context=None):
        samples = self.sample(num_samples, context=context)
        if context is not None:
            raise NotImplementedError("oracle shim: no context support here")
        log_prob = self.log_prob(samples, context=context)
        return samples, log_prob

    def mean(self, context=None):
        if context is not None:
            context = torch.as_tensor(context)
        return self._mean(context)

    def _mean(self, context):
        raise NoMeanException()
