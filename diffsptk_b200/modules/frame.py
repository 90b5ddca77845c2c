"""Frame: overlapping frames of a waveform (drop-in for diffsptk/modules/frame.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import filter_values, pad_mode_id
from .base import BaseFunctionalModule, Precomputed


class Frame(BaseFunctionalModule):
    """``(..., T) -> (..., (T-1)//P+1, L)``; CUDA kernel ``dsb200_frame`` (bit-exact gather unless zmean)."""

    def __init__(self, frame_length: int, frame_period: int, *, center: bool = True, zmean: bool = False,
                 mode: str = "constant") -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return Frame._apply_precomputed(Frame._precompute(*args, **kwargs), x=x)

    @staticmethod
    def _check(frame_length: int, frame_period: int) -> None:
        if frame_length <= 0:
            raise ValueError("frame_length must be positive.")
        if frame_period <= 0:
            raise ValueError("frame_period must be positive.")

    @staticmethod
    def _precompute(frame_length: int, frame_period: int, center: bool = True, zmean: bool = False,
                    mode: str = "constant") -> Precomputed:
        Frame._check(frame_length, frame_period)
        pad_mode_id(mode)
        return Precomputed(values=dict(frame_length=frame_length, frame_period=frame_period, center=center,
                                       zmean=zmean, mode=mode))

    @staticmethod
    def _forward(x: torch.Tensor, *, frame_length: int, frame_period: int, center: bool, zmean: bool,
                 mode: str) -> torch.Tensor:
        return ops.frame(x, frame_length, frame_period, center, zmean, pad_mode_id(mode))
