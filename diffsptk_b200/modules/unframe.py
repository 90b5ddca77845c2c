"""Windowed overlap-add (drop-in for diffsptk/modules/unframe.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed


class Unframe(BaseFunctionalModule):
    """``(..., T/P, L) -> (..., T)``; kernel ``dsb200_unframe``.  Buffer name ``window`` as in the reference."""

    _takes_input_size = True

    def __init__(self, frame_length: int, frame_period: int, *, center: bool = True,
                 window: str | int = "rectangular", norm: str | int = "none", symmetric: bool = True,
                 learnable: bool = False, device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = frame_length
        self._register_precomputed(self._precompute(**filter_values(locals(), drop_keys=["learnable"])),
                                   learnable=learnable)

    def forward(self, y: torch.Tensor, out_length: int | None = None) -> torch.Tensor:
        check_size(y.size(-1), self.in_dim, "length of waveform")
        return self._call_forward(y, out_length)

    @staticmethod
    def _func(y: torch.Tensor, out_length: int | None, *args, **kwargs) -> torch.Tensor:
        pre = Unframe._precompute(y.size(-1), *args, **kwargs, device=y.device, dtype=y.dtype)
        return Unframe._apply_precomputed(pre, y=y, out_length=out_length)

    @staticmethod
    def _check(frame_length: int, frame_period: int) -> None:
        if frame_length <= 0:
            raise ValueError("frame_length must be positive.")
        if frame_length < frame_period:
            raise ValueError("frame_period must be less than or equal to frame_length.")
        if frame_period <= 0:
            raise ValueError("frame_period must be positive.")

    @staticmethod
    def _precompute(frame_length: int, frame_period: int, center: bool = True, window: str | int = "rectangular",
                    norm: str | int = "none", symmetric: bool = True, device: torch.device | None = None,
                    dtype: torch.dtype | None = None) -> Precomputed:
        Unframe._check(frame_length, frame_period)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        table = tables.make_window(frame_length, window, norm, symmetric, device=device, dtype=dtype)
        # same shape as the reference's buffer / parameter (1, L, 1) (unframe.py:150-152): state_dicts interchange
        return Precomputed(values=dict(frame_period=frame_period, center=center),
                           tensors={"window": table.view(1, -1, 1)})

    @staticmethod
    def _forward(y: torch.Tensor, out_length: int | None, *, frame_period: int, center: bool,
                 window: torch.Tensor) -> torch.Tensor:
        if y.dim() <= 1:
            raise ValueError("Input must be at least 2D tensor.")
        T = ops.unframe_length(y.size(-2), y.size(-1), frame_period, center, out_length)
        return ops.unframe(y, window.reshape(-1), T, frame_period, center)   # differentiable in y and in the window
