"""Gain normalisation of generalized cepstra (drop-in for diffsptk/modules/gnorm.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import check_size, filter_values, get_gamma
from .base import BaseFunctionalModule, Precomputed


class GeneralizedCepstrumGainNormalization(BaseFunctionalModule):
    """``(..., M+1) -> (..., M+1)``: ``K = (1 + gamma c_0)^(1/gamma)`` (``exp c_0`` for gamma 0),
    ``c_m / (1 + gamma c_0)`` (gnorm.py:101-112); kernel ``dsb200_rowconv``."""

    _takes_input_size = True

    def __init__(self, cep_order: int, gamma: float = 0, c: int | None = None) -> None:
        super().__init__()
        self.in_dim = cep_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        check_size(x.size(-1), self.in_dim, "dimension of cepstrum")
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = GeneralizedCepstrumGainNormalization._precompute(x.size(-1) - 1, *args, **kwargs)
        return GeneralizedCepstrumGainNormalization._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(cep_order: int, gamma: float, c: int | None) -> None:
        if cep_order < 0:
            raise ValueError("cep_order must be non-negative.")
        if 1 < abs(gamma):
            raise ValueError("gamma must be in [-1, 1].")
        if c is not None and c < 1:
            raise ValueError("c must be greater than or equal to 1.")

    @staticmethod
    def _precompute(cep_order: int, gamma: float, c: int | None = None) -> Precomputed:
        GeneralizedCepstrumGainNormalization._check(cep_order, gamma, c)
        return Precomputed(values={"gamma": get_gamma(gamma, c)})

    @staticmethod
    def _forward(x: torch.Tensor, *, gamma: float) -> torch.Tensor:
        return ops.rowconv(x, ops.CONV_GNORM, gamma)
