"""Logit / Sigmoid (oracle only). Serves flows/utils.py:344-345 (pre_transform)."""

import torch
from torch.nn import functional as F

from ..utils import torchutils
from .base import InputOutsideDomain, Transform


class Sigmoid(Transform):
    def __init__(self, temperature=1, eps=1e-6, learn_temperature=False):
        super().__init__()
        self.eps = eps
        if learn_temperature:
            self.temperature = torch.nn.Parameter(torch.Tensor([temperature]))
        else:
            self.temperature = torch.Tensor([temperature])

    def forward(self, inputs, context=None):
        inputs = self.temperature * inputs
        outputs = torch.sigmoid(inputs)
        logabsdet = torchutils.sum_except_batch(
            torch.log(self.temperature) - F.softplus(-inputs) - F.softplus(inputs)
        )
        return outputs, logabsdet

    def inverse(self, inputs, context=None):
        if torch.min(inputs) < 0 or torch.max(inputs) > 1:
            raise InputOutsideDomain()
        inputs = torch.clamp(inputs, self.eps, 1 - self.eps)
        outputs = (1 / self.temperature) * (
            torch.log(inputs) - torch.log1p(-inputs)
        )
        logabsdet = -torchutils.sum_except_batch(
            torch.log(self.temperature)
            - F.softplus(-self.temperature * outputs)
            - F.softplus(self.temperature * outputs)
        )
        return outputs, logabsdet


class Logit(Transform):
    """Inverse of Sigmoid."""

    def __init__(self, temperature=1, eps=1e-6):
        super().__init__()
        self._sigmoid = Sigmoid(temperature=temperature, eps=eps)

    def forward(self, inputs, context=None):
        return self._sigmoid.inverse(inputs, context)

    def inverse(self, inputs, context=None):
        return self._sigmoid(inputs, context)
