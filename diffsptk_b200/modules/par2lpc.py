"""PARCOR -> LPC coefficients (drop-in for diffsptk/modules/par2lpc.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed
from .lpc2par import LinearPredictiveCoefficientsToParcorCoefficients


class ParcorCoefficientsToLinearPredictiveCoefficients(BaseFunctionalModule):
    """``(..., M+1) -> (..., M+1)``: ``[K, k_1..k_M] -> [K, a_1..a_M] / gamma`` by the step-up recursion
    (par2lpc.py:100-107); kernel ``dsb200_rowconv``."""

    _takes_input_size = True

    def __init__(self, lpc_order: int, gamma: float = 1, c: int | None = None) -> None:
        super().__init__()
        self.in_dim = lpc_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, k: torch.Tensor) -> torch.Tensor:
        check_size(k.size(-1), self.in_dim, "dimension of PARCOR")
        return self._call_forward(k)

    @staticmethod
    def _func(k: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = ParcorCoefficientsToLinearPredictiveCoefficients._precompute(k.size(-1) - 1, *args, **kwargs)
        return ParcorCoefficientsToLinearPredictiveCoefficients._apply_precomputed(pre, k=k)

    @staticmethod
    def _check(*args, **kwargs) -> None:
        raise NotImplementedError

    @staticmethod
    def _precompute(*args, **kwargs) -> Precomputed:
        return LinearPredictiveCoefficientsToParcorCoefficients._precompute(*args, **kwargs)

    @staticmethod
    def _forward(k: torch.Tensor, *, gamma: float) -> torch.Tensor:
        return ops.rowconv(k, ops.CONV_PAR2LPC, gamma)
