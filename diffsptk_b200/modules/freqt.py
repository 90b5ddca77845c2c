"""Frequency transform of cepstra (drop-in for diffsptk/modules/freqt.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed


def _warp_check(in_order: int, out_order: int, alpha: float) -> None:
    if in_order < 0:
        raise ValueError("in_order must be non-negative.")
    if out_order < 0:
        raise ValueError("out_order must be non-negative.")
    if 1 <= abs(alpha):
        raise ValueError("alpha must be in (-1, 1).")


class FrequencyTransform(BaseFunctionalModule):
    """``(..., M1+1) -> (..., M2+1)`` = ``c @ A``; kernel ``dsb200_rowmat``; buffer ``A`` as in the reference."""

    _takes_input_size = True

    def __init__(self, in_order: int, out_order: int, alpha: float = 0, device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = in_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, c: torch.Tensor) -> torch.Tensor:
        check_size(c.size(-1), self.in_dim, "dimension of cepstrum")
        return self._call_forward(c)

    @staticmethod
    def _func(c: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = FrequencyTransform._precompute(c.size(-1) - 1, *args, **kwargs, device=c.device, dtype=c.dtype)
        return FrequencyTransform._apply_precomputed(pre, c=c)

    @staticmethod
    def _check(in_order: int, out_order: int, alpha: float) -> None:
        _warp_check(in_order, out_order, alpha)

    @staticmethod
    def _precompute(in_order: int, out_order: int, alpha: float, device: torch.device | None,
                    dtype: torch.dtype | None) -> Precomputed:
        FrequencyTransform._check(in_order, out_order, alpha)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        return Precomputed(tensors={"A": tables.make_freqt_matrix(in_order, out_order, alpha, device, dtype)})

    @staticmethod
    def _forward(c: torch.Tensor, *, A: torch.Tensor) -> torch.Tensor:
        return ops.rowmat(c, A)
