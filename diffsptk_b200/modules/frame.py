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
        if x.dtype in (torch.float32, torch.float64):
            return ops.frame(x, frame_length, frame_period, center, zmean, pad_mode_id(mode))
        # The reference's Frame is pad + unfold, which keeps the input dtype (half / bfloat16 / integer waveforms stay
        # what they are, frame.py:130-141).  The gather is exact in a wider float type, so run it there and cast back:
        # float32 holds half / bfloat16 / int8 / int16 exactly, float64 holds int32 (and int64 up to 2^53).  With zmean
        # the reference's mean of a half tensor is a half computation; the wider result rounded back is at least as close.
        wide = torch.float32 if (x.dtype.is_floating_point or x.element_size() <= 2) else torch.float64
        y = ops.frame(x.to(wide), frame_length, frame_period, center, zmean, pad_mode_id(mode))
        if zmean and not x.dtype.is_floating_point:
            return y.to(torch.float32)          # the reference raises for integer zmean (mean of a Long tensor)
        return y.to(x.dtype)
