"""MLSA filter coefficients -> mel-cepstrum (drop-in for diffsptk/modules/b2mc.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed


class MLSADigitalFilterCoefficientsToMelCepstrum(BaseFunctionalModule):
    """``(..., M+1) -> (..., M+1)``: ``mc_m = b_m + alpha b_{m+1}`` as ``b @ A`` (b2mc.py:104-119);
    kernel ``dsb200_rowmat``; buffer ``A`` as in the reference."""

    _takes_input_size = True

    def __init__(self, cep_order: int, alpha: float = 0, device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = cep_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, b: torch.Tensor) -> torch.Tensor:
        check_size(b.size(-1), self.in_dim, "dimension of cepstrum")
        return self._call_forward(b)

    @staticmethod
    def _func(b: torch.Tensor, alpha: float) -> torch.Tensor:
        pre = MLSADigitalFilterCoefficientsToMelCepstrum._precompute(b.size(-1) - 1, alpha, device=b.device,
                                                                     dtype=b.dtype)
        return MLSADigitalFilterCoefficientsToMelCepstrum._apply_precomputed(pre, b=b)

    @staticmethod
    def _check(cep_order: int, alpha: float) -> None:
        if cep_order < 0:
            raise ValueError("cep_order must be non-negative.")
        if 1 <= abs(alpha):
            raise ValueError("alpha must be in (-1, 1).")

    @staticmethod
    def _precompute(cep_order: int, alpha: float, device: torch.device | None,
                    dtype: torch.dtype | None) -> Precomputed:
        MLSADigitalFilterCoefficientsToMelCepstrum._check(cep_order, alpha)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        return Precomputed(tensors={"A": tables.make_b2mc_matrix(cep_order, alpha, device, dtype)})

    @staticmethod
    def _forward(b: torch.Tensor, *, A: torch.Tensor) -> torch.Tensor:
        return ops.rowmat(b, A)
