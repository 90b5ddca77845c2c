"""Window: multiply frames by a window table and zero-pad (drop-in for diffsptk/modules/window.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed


class Window(BaseFunctionalModule):
    """``(..., L1) -> (..., L2)``.  Buffer name ``window`` as in the reference (window.py:181-183)."""

    _takes_input_size = True

    def __init__(self, in_length: int, out_length: int | None = None, *, window: str | int = "blackman",
                 norm: str | int = "power", symmetric: bool = True, learnable: bool = False,
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = in_length
        self._register_precomputed(self._precompute(**filter_values(locals(), drop_keys=["learnable"])),
                                   learnable=learnable)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        check_size(x.size(-1), self.in_dim, "input length")
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = Window._precompute(x.size(-1), *args, **kwargs, device=x.device, dtype=x.dtype)
        return Window._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(in_length: int, out_length: int | None) -> None:
        if in_length <= 0:
            raise ValueError("in_length must be positive.")
        if out_length is not None and out_length <= 0:
            raise ValueError("out_length must be positive.")

    @staticmethod
    def _precompute(in_length: int, out_length: int | None, window: str | int, norm: str | int,
                    symmetric: bool, device: torch.device | None, dtype: torch.dtype | None) -> Precomputed:
        Window._check(in_length, out_length)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None  # integer waveforms: default-dtype table, result promotes like x * window
        table = tables.make_window(in_length, window, norm, symmetric, device=device, dtype=dtype)
        return Precomputed(values={"out_length": out_length}, tensors={"window": table})

    @staticmethod
    def _forward(x: torch.Tensor, *, out_length: int | None, window: torch.Tensor) -> torch.Tensor:
        return ops.window(x, window, x.size(-1) if out_length is None else out_length)
