"""The element-wise stage of an affine coupling layer on its own.

``glasflow.nflows``' ``AffineCouplingTransform._coupling_transform_forward`` /
``_inverse`` (what /root/reference/src/nessai/flows/realnvp.py:110-112 instantiates)
with the conditioner output supplied: the HBM-class part of the flow.
In the product the same arithmetic is fused into the flow kernels' epilogues; this
entry point exists for callers that bring their own conditioner and as the kernel
the "coupling forward vs HBM roofline" figure is measured on.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def coupling_transform(x: torch.Tensor, params: torch.Tensor, transform_features, additive=False,
                       inverse=False):
    """``x`` (n, D) fp32 CUDA, ``params`` (n, 2*d_tr) = shift | unconstrained scale
    (``(n, d_tr)`` when ``additive``).  Returns ``(y (n, D), logabsdet (n,))``."""
    if x.device.type != "cuda":
        raise RuntimeError("nessai_b200 kernels run on a CUDA device; there is no CPU fallback")
    x = x.to(torch.float32).contiguous()
    params = params.to(device=x.device, dtype=torch.float32).contiguous()
    tf = np.ascontiguousarray(np.asarray(transform_features, dtype=np.int32))
    n, D = x.shape
    d_tr = int(tf.size)
    if params.shape != (n, d_tr * (1 if additive else 2)):
        raise ValueError(f"params has shape {tuple(params.shape)}, expected {(n, d_tr * (1 if additive else 2))}")
    y = torch.empty_like(x)
    ld = torch.empty(n, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(
            _lib.load().nb200_coupling_transform(
                C.c_void_p(x.data_ptr()), C.c_void_p(params.data_ptr()), C.c_void_p(y.data_ptr()),
                C.c_void_p(ld.data_ptr()), n, D, tf.ctypes.data_as(C.c_void_p), d_tr,
                int(bool(additive)), int(bool(inverse)),
                C.c_void_p(torch.cuda.current_stream().cuda_stream),
            ),
            "nb200_coupling_transform",
        )
    return y, ld
