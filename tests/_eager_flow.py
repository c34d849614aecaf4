"""TEST INFRASTRUCTURE: differentiable train-mode flow on the flat parameter buffer
(torch ops + autograd), the on-device gradient cross-check for the fused training
kernels (``nessai_b200/csrc/train.cuh``).  Restates the train-mode arithmetic the
reference delegates to glasflow.nflows: batch-statistics BatchNorm,
uncached LU, affine coupling, MLP / ResidualNet conditioner.  The product never
imports this file.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from nessai_b200.spec import ACT_RELU, ACT_TANH, FlowSpec


def _act(kind):
    if kind == ACT_RELU:
        return F.relu
    if kind == ACT_TANH:
        return torch.tanh
    return lambda x: x * torch.sigmoid(x)


class EagerFlow:
    """Functional view of a :class:`FlowSpec` over a flat ``theta`` tensor."""

    def __init__(self, spec: FlowSpec, ints: dict, device):
        if spec.ftype != "realnvp":
            raise NotImplementedError(
                "nessai_b200: training is implemented for RealNVP flows only"
            )
        self.spec = spec
        self.device = torch.device(device)
        self.act = _act(spec.activation)
        D = spec.D
        self._tril = torch.tril_indices(D, D, -1, device=self.device)
        self._triu = torch.triu_indices(D, D, 1, device=self.device)
        self._eye = torch.eye(D, device=self.device)
        self.update_ints(ints)

    def update_ints(self, ints):
        dev = self.device
        self.perm = {
            ls.perm_key: torch.as_tensor(ints[ls.perm_key], device=dev)
            for ls in self.spec.layers
            if ls.perm_key is not None
        }
        self.idf = [torch.as_tensor(ls.identity, device=dev) for ls in self.spec.layers]
        self.trf = [torch.as_tensor(ls.transform, device=dev) for ls in self.spec.layers]

    def _get(self, theta, key):
        """``theta`` = (params, float_buffers): two flat tensors."""
        e = self.spec.by_key[key]
        if e.kind == "param":
            return theta[0][e.offset : e.offset + e.size].view(e.shape)
        off = e.offset - self.spec.n_params
        return theta[1][off : off + e.size].view(e.shape)

    def _net(self, theta, ls, h):
        sp, act = self.spec, self.act
        lins = ls.linears

        def lin(i, v):
            return F.linear(v, self._get(theta, lins[i].weight), self._get(theta, lins[i].bias))

        if sp.net == "mlp":
            for i in range(len(lins) - 1):
                h = act(lin(i, h))
            return lin(len(lins) - 1, h)
        h = lin(0, h)
        for b in range(sp.n_layers):
            t = lin(1 + 2 * b, act(h))
            t = lin(2 + 2 * b, act(t))
            h = h + t
        return lin(len(lins) - 1, h)

    def log_prob(self, theta, x, training: bool, update_running: bool = True):
        """``NFlow.log_prob`` (/root/reference/src/nessai/flows/base.py:236-245).

        ``theta`` = ``(params, float_buffers)``: the two flat fp32 tensors of the
        flow; in training mode the BatchNorm running statistics inside
        ``float_buffers`` are EMA-updated in place (no grad), exactly like
        nflows' BatchNorm.
        """
        sp = self.spec
        D = sp.D
        n = x.shape[0]
        ld = x.new_zeros(n)
        h = x
        for i, ls in enumerate(sp.layers):
            if ls.perm_key is not None:
                h = h[:, self.perm[ls.perm_key]]
            if ls.lu_prefix is not None:
                lo = self._eye.clone()
                lo[self._tril[0], self._tril[1]] = self._get(theta, f"{ls.lu_prefix}.lower_entries")
                diag = F.softplus(self._get(theta, f"{ls.lu_prefix}.unconstrained_upper_diag")) + sp.LU_EPS
                up = torch.diag(diag)
                up[self._triu[0], self._triu[1]] = self._get(theta, f"{ls.lu_prefix}.upper_entries")
                h = F.linear(F.linear(h, up), lo, self._get(theta, f"{ls.lu_prefix}.bias"))
                ld = ld + torch.sum(torch.log(diag))
            ident = h[:, self.idf[i]]
            tr = h[:, self.trf[i]]
            params = self._net(theta, ls, ident)
            d_tr = tr.shape[1]
            if sp.volume_preserving:
                tr = tr + params
            else:
                shift = params[:, :d_tr]
                scale = torch.sigmoid(params[:, d_tr:] + 2) + 1e-3
                tr = tr * scale + shift
                ld = ld + torch.log(scale).sum(1)
            out = torch.empty_like(h)
            out[:, self.idf[i]] = ident
            out[:, self.trf[i]] = tr
            h = out
            if ls.bn_prefix is not None:
                w = F.softplus(self._get(theta, f"{ls.bn_prefix}.unconstrained_weight")) + sp.BN_EPS
                beta = self._get(theta, f"{ls.bn_prefix}.bias")
                rm = self._get(theta, f"{ls.bn_prefix}.running_mean")
                rv = self._get(theta, f"{ls.bn_prefix}.running_var")
                if training:
                    mean, var = h.mean(0), h.var(0)
                    if update_running:
                        with torch.no_grad():
                            rm.mul_(1 - sp.BN_MOMENTUM).add_(mean.detach() * sp.BN_MOMENTUM)
                            rv.mul_(1 - sp.BN_MOMENTUM).add_(var.detach() * sp.BN_MOMENTUM)
                else:
                    mean, var = rm.detach(), rv.detach()
                h = w * ((h - mean) / torch.sqrt(var + sp.BN_EPS)) + beta
                ld = ld + torch.sum(torch.log(w) - 0.5 * torch.log(var + sp.BN_EPS))
        base = -0.5 * torch.sum(h * h, dim=1) - 0.5 * D * math.log(2 * math.pi)
        return base + ld
