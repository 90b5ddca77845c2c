"""All-pole -> all-zero filter coefficients (drop-in for diffsptk/modules/norm0.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed


class AllPoleToAllZeroDigitalFilterCoefficients(BaseFunctionalModule):
    """``(..., M+1) -> (..., M+1)``: ``[K, a_1..a_M] -> [1/K, a_1/K..a_M/K]`` (norm0.py:88-94);
    kernel ``dsb200_rowconv``."""

    _takes_input_size = True

    def __init__(self, filter_order: int) -> None:
        super().__init__()
        self.in_dim = filter_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, a: torch.Tensor) -> torch.Tensor:
        check_size(a.size(-1), self.in_dim, "dimension of coefficients")
        return self._forward(a)

    @staticmethod
    def _func(a: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        AllPoleToAllZeroDigitalFilterCoefficients._precompute(a.size(-1) - 1, *args, **kwargs)
        return AllPoleToAllZeroDigitalFilterCoefficients._forward(a)

    @staticmethod
    def _check(filter_order: int) -> None:
        if filter_order < 0:
            raise ValueError("filter_order must be non-negative.")

    @staticmethod
    def _precompute(filter_order: int) -> Precomputed:
        AllPoleToAllZeroDigitalFilterCoefficients._check(filter_order)
        return Precomputed()

    @staticmethod
    def _forward(a: torch.Tensor) -> torch.Tensor:
        return ops.rowconv(a, ops.CONV_NORM0, 0.0)
