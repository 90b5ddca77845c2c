"""LPC -> PARCOR coefficients (drop-in for diffsptk/modules/lpc2par.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import check_size, filter_values, get_gamma
from .base import BaseFunctionalModule, Precomputed


class LinearPredictiveCoefficientsToParcorCoefficients(BaseFunctionalModule):
    """``(..., M+1) -> (..., M+1)``: ``[K, a_1..a_M] -> [K, k_1..k_M]`` by the step-down recursion.

    Kernel ``dsb200_rowconv`` (csrc/convert.cu): one thread per row on a shared-memory tile, rows read and written
    once with coalesced accesses.  The reference runs M dependent tensor ops with a flip each (lpc2par.py:104-120).
    """

    _takes_input_size = True

    def __init__(self, lpc_order: int, gamma: float = 1, c: int | None = None) -> None:
        super().__init__()
        self.in_dim = lpc_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, a: torch.Tensor) -> torch.Tensor:
        check_size(a.size(-1), self.in_dim, "dimension of LPC")
        return self._call_forward(a)

    @staticmethod
    def _func(a: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = LinearPredictiveCoefficientsToParcorCoefficients._precompute(a.size(-1) - 1, *args, **kwargs)
        return LinearPredictiveCoefficientsToParcorCoefficients._apply_precomputed(pre, a=a)

    @staticmethod
    def _check(lpc_order: int, gamma: float, c: int | None) -> None:
        if lpc_order < 0:
            raise ValueError("lpc_order must be non-negative.")
        if 1 < abs(gamma):
            raise ValueError("gamma must be in [-1, 1].")
        if c is not None and c < 1:
            raise ValueError("c must be greater than or equal to 1.")

    @staticmethod
    def _precompute(lpc_order: int, gamma: float = 1, c: int | None = None) -> Precomputed:
        LinearPredictiveCoefficientsToParcorCoefficients._check(lpc_order, gamma, c)
        return Precomputed(values={"gamma": get_gamma(gamma, c)})

    @staticmethod
    def _forward(a: torch.Tensor, *, gamma: float) -> torch.Tensor:
        return ops.rowconv(a, ops.CONV_LPC2PAR, gamma)
