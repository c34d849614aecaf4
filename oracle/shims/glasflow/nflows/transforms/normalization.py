"""Normalisation transforms (oracle only).

Serves /root/reference/src/nessai/flows/realnvp.py:186-189,205-206 and the
reset constants pinned by utils.py:262-272 /
/root/reference/tests/test_flows/test_flow_utils.py:158-173.

NOTE (unpinned): upstream nflows registers ``running_var`` as zeros; the
reference's ``reset_weights`` sets it to one.  The initial value only enters
through the 0.9**k EMA tail.  We follow upstream (zeros).
"""

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .base import InverseNotAvailable, Transform


class BatchNorm(Transform):
    """Flow-layer batch norm: y = w (x - mean)/sqrt(var + eps) + b."""

    def __init__(self, features, eps=1e-5, momentum=0.1, affine=True):
        if not isinstance(features, int) or features < 1:
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.momentum = momentum
        self.eps = eps
        constant = np.log(np.exp(1 - eps) - 1)
        self.unconstrained_weight = nn.Parameter(constant * torch.ones(features))
        self.bias = nn.Parameter(torch.zeros(features))
        self.register_buffer("running_mean", torch.zeros(features))
        self.register_buffer("running_var", torch.zeros(features))

    @property
    def weight(self):
        return F.softplus(self.unconstrained_weight) + self.eps

    def forward(self, inputs, context=None):
        if inputs.dim() != 2:
            raise ValueError(
                "Expected 2-dim inputs, got inputs of shape: {}".format(
                    inputs.shape
                )
            )
        if self.training:
            mean, var = inputs.mean(0), inputs.var(0)
            self.running_mean.mul_(1 - self.momentum).add_(
                mean.detach() * self.momentum
            )
            self.running_var.mul_(1 - self.momentum).add_(
                var.detach() * self.momentum
            )
        else:
            mean, var = self.running_mean, self.running_var

        outputs = (
            self.weight * ((inputs - mean) / torch.sqrt((var + self.eps)))
            + self.bias
        )
        logabsdet_ = torch.log(self.weight) - 0.5 * torch.log(var + self.eps)
        logabsdet = torch.sum(logabsdet_) * inputs.new_ones(inputs.shape[0])
        return outputs, logabsdet

    def inverse(self, inputs, context=None):
        if self.training:
            raise InverseNotAvailable(
                "Batch norm inverse is only available in eval mode, not in "
                "training mode."
            )
        if inputs.dim() != 2:
            raise ValueError(
                "Expected 2-dim inputs, got inputs of shape: {}".format(
                    inputs.shape
                )
            )
        outputs = (
            torch.sqrt(self.running_var + self.eps)
            * ((inputs - self.bias) / self.weight)
            + self.running_mean
        )
        logabsdet_ = -torch.log(self.weight) + 0.5 * torch.log(
            self.running_var + self.eps
        )
        logabsdet = torch.sum(logabsdet_) * inputs.new_ones(inputs.shape[0])
        return outputs, logabsdet


class ActNorm(Transform):
    """Activation normalisation with data-dependent init (Glow)."""

    def __init__(self, features):
        if not isinstance(features, int) or features < 1:
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.register_buffer("initialized", torch.tensor(False, dtype=torch.bool))
        self.log_scale = nn.Parameter(torch.zeros(features))
        self.shift = nn.Parameter(torch.zeros(features))

    @property
    def scale(self):
        return torch.exp(self.log_scale)

    def forward(self, inputs, context=None):
        if inputs.dim() != 2:
            raise ValueError("Expecting inputs to be a 2D tensor.")
        if self.training and not self.initialized:
            self._initialize(inputs)
        scale = self.scale.view(1, -1)
        shift = self.shift.view(1, -1)
        outputs = scale * inputs + shift
        logabsdet = torch.sum(self.log_scale) * outputs.new_ones(inputs.shape[0])
        return outputs, logabsdet

    def inverse(self, inputs, context=None):
        if inputs.dim() != 2:
            raise ValueError("Expecting inputs to be a 2D tensor.")
        scale = self.scale.view(1, -1)
        shift = self.shift.view(1, -1)
        outputs = (inputs - shift) / scale
        logabsdet = -torch.sum(self.log_scale) * outputs.new_ones(inputs.shape[0])
        return outputs, logabsdet

    def _initialize(self, inputs):
        with torch.no_grad():
            std = inputs.std(dim=0)
            mu = (inputs / std).mean(dim=0)
            self.log_scale.data = -torch.log(std)
            self.shift.data = -mu
            self.initialized.data = torch.tensor(True, dtype=torch.bool)
