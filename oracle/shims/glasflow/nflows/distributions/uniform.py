"""BoxUniform (oracle only). Serves nessai/utils/distributions.py:8."""

from typing import Union

import torch
from torch import distributions


class BoxUniform(distributions.Independent):
    def __init__(
        self,
        low: Union[torch.Tensor, float],
        high: Union[torch.Tensor, float],
        reinterpreted_batch_ndims: int = 1,
    ):
        super().__init__(
            distributions.Uniform(low=low, high=high), reinterpreted_batch_ndims
        )
