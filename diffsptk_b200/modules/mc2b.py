"""Mel-cepstrum -> MLSA filter coefficients (drop-in for diffsptk/modules/mc2b.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import check_size, filter_values
from .b2mc import MLSADigitalFilterCoefficientsToMelCepstrum
from .base import BaseFunctionalModule, Precomputed


class MelCepstrumToMLSADigitalFilterCoefficients(BaseFunctionalModule):
    """``(..., M+1) -> (..., M+1)``: ``b_m = mc_m - alpha b_{m+1}`` as ``mc @ A`` (mc2b.py:107-122);
    kernel ``dsb200_rowmat``; buffer ``A`` as in the reference."""

    _takes_input_size = True

    def __init__(self, cep_order: int, alpha: float = 0, device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = cep_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, mc: torch.Tensor) -> torch.Tensor:
        check_size(mc.size(-1), self.in_dim, "dimension of cepstrum")
        return self._call_forward(mc)

    @staticmethod
    def _func(mc: torch.Tensor, alpha: float) -> torch.Tensor:
        pre = MelCepstrumToMLSADigitalFilterCoefficients._precompute(mc.size(-1) - 1, alpha, device=mc.device,
                                                                     dtype=mc.dtype)
        return MelCepstrumToMLSADigitalFilterCoefficients._apply_precomputed(pre, mc=mc)

    @staticmethod
    def _check(*args, **kwargs) -> None:
        MLSADigitalFilterCoefficientsToMelCepstrum._check(*args, **kwargs)

    @staticmethod
    def _precompute(cep_order: int, alpha: float, device: torch.device | None,
                    dtype: torch.dtype | None) -> Precomputed:
        MelCepstrumToMLSADigitalFilterCoefficients._check(cep_order, alpha)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        return Precomputed(tensors={"A": tables.make_mc2b_matrix(cep_order, alpha, device, dtype)})

    @staticmethod
    def _forward(mc: torch.Tensor, *, A: torch.Tensor) -> torch.Tensor:
        return ops.rowmat(mc, A)
