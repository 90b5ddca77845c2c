"""Inverse gain normalisation of generalized cepstra (drop-in for diffsptk/modules/ignorm.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed
from .gnorm import GeneralizedCepstrumGainNormalization


class GeneralizedCepstrumInverseGainNormalization(BaseFunctionalModule):
    """``(..., M+1) -> (..., M+1)``: ``c_0 = (K^gamma - 1) / gamma`` (``log K`` for gamma 0), ``c_m K^gamma``
    (ignorm.py:98-109); kernel ``dsb200_rowconv``."""

    _takes_input_size = True

    def __init__(self, cep_order: int, gamma: float = 0, c: int | None = None) -> None:
        super().__init__()
        self.in_dim = cep_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, y: torch.Tensor) -> torch.Tensor:
        check_size(y.size(-1), self.in_dim, "dimension of cepstrum")
        return self._call_forward(y)

    @staticmethod
    def _func(y: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = GeneralizedCepstrumInverseGainNormalization._precompute(y.size(-1) - 1, *args, **kwargs)
        return GeneralizedCepstrumInverseGainNormalization._apply_precomputed(pre, y=y)

    @staticmethod
    def _check(*args, **kwargs) -> None:
        raise NotImplementedError

    @staticmethod
    def _precompute(*args, **kwargs) -> Precomputed:
        return GeneralizedCepstrumGainNormalization._precompute(*args, **kwargs)

    @staticmethod
    def _forward(y: torch.Tensor, *, gamma: float) -> torch.Tensor:
        return ops.rowconv(y, ops.CONV_IGNORM, gamma)
