"""nflows.utils.torchutils restatement (oracle only)."""
import torch


def sum_except_batch(x, num_batch_dims=1):
    """Sum over every dimension but the leading ``num_batch_dims``."""
    reduce_dims = list(range(num_batch_dims, x.ndimension()))
    if not reduce_dims:
        return x
    return torch.sum(x, dim=reduce_dims)


def tile(x, n):
    """Repeat every element ``n`` times consecutively: [a,a,..,b,b,..]."""
    x_ = x.reshape(-1)
    x_ = x_.repeat(n)
    x_ = x_.reshape(n, -1)
    x_ = x_.transpose(1, 0)
    return x_.reshape(-1)


def searchsorted(bin_locations, inputs, eps=1e-6):
    """Bin index = #(knots <= x) - 1, with the last knot nudged by eps."""
    bin_locations[..., -1] += eps
    return torch.sum(inputs[..., None] >= bin_locations, dim=-1) - 1
