"""LPC analysis = levdur(acorr(x)), fused (drop-in for diffsptk/modules/lpc.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import filter_values, get_layer
from .acorr import Autocorrelation
from .base import BaseFunctionalModule, Precomputed
from .levdur import LevinsonDurbin, default_eps


class LinearPredictiveCodingAnalysis(BaseFunctionalModule):
    """``(..., L) -> (..., M+1)``; kernel ``dsb200_lpc`` (lag sums + recursion, no intermediate tensor)."""

    _takes_input_size = True

    def __init__(self, frame_length: int, lpc_order: int, eps: float | None = None,
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = LinearPredictiveCodingAnalysis._precompute(x.size(-1), *args, **kwargs, device=x.device,
                                                         dtype=x.dtype, module=False)
        return LinearPredictiveCodingAnalysis._apply_precomputed(pre, x=x)

    @staticmethod
    def _check() -> None:
        pass

    @staticmethod
    def _precompute(frame_length: int, lpc_order: int, eps: float | None, device: torch.device | None,
                    dtype: torch.dtype | None, module: bool = True) -> Precomputed:
        LinearPredictiveCodingAnalysis._check()
        Autocorrelation._check(frame_length, lpc_order)
        LevinsonDurbin._check(lpc_order, eps)
        fdtype = dtype if (dtype is None or dtype.is_floating_point) else None
        values = dict(frame_length=frame_length, lpc_order=lpc_order, eps=float(default_eps(eps, fdtype)))
        if not module:
            return Precomputed(values=values)
        acorr = get_layer(True, Autocorrelation, dict(frame_length=frame_length, acr_order=lpc_order))
        levdur = get_layer(True, LevinsonDurbin, dict(lpc_order=lpc_order, eps=eps, device=device, dtype=dtype))
        return Precomputed(values=values, layers={"acorr": acorr, "levdur": levdur})

    @staticmethod
    def _forward(x: torch.Tensor, *, frame_length: int, lpc_order: int, eps: float, acorr=None,
                 levdur=None) -> torch.Tensor:
        if x.size(-1) != frame_length:
            raise ValueError(f"Unexpected length of waveform (input {x.size(-1)} vs target {frame_length}).")
        return ops.lpc(x, lpc_order, eps)
