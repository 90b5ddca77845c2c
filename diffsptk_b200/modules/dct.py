"""Discrete cosine transform (drop-in for diffsptk/modules/dct.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed


class DiscreteCosineTransform(BaseFunctionalModule):
    """``(..., L) -> (..., L)`` = ``x @ W``; kernel ``dsb200_rowmat``; buffer ``W`` as in the reference."""

    _takes_input_size = True

    def __init__(self, dct_length: int, dct_type: int = 2, device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = dct_length
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        check_size(x.size(-1), self.in_dim, "dimension of input")
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = DiscreteCosineTransform._precompute(x.size(-1), *args, **kwargs, device=x.device, dtype=x.dtype)
        return DiscreteCosineTransform._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(dct_length: int, dct_type: int) -> None:
        if dct_length <= 0:
            raise ValueError("dct_length must be positive.")
        if not 1 <= dct_type <= 4:
            raise ValueError("dct_type must be in [1, 4].")

    @staticmethod
    def _precompute(dct_length: int, dct_type: int, device: torch.device | None,
                    dtype: torch.dtype | None) -> Precomputed:
        DiscreteCosineTransform._check(dct_length, dct_type)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        return Precomputed(tensors={"W": tables.make_dct_matrix(dct_length, dct_type, device, dtype)})

    @staticmethod
    def _forward(x: torch.Tensor, *, W: torch.Tensor) -> torch.Tensor:
        return ops.rowmat(x, W)
