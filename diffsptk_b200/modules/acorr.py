"""Autocorrelation (drop-in for diffsptk/modules/acorr.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed

_FORMATS = {"naive": 0, "normalized": 1, "biased": 2, "unbiased": 3}


class Autocorrelation(BaseFunctionalModule):
    """``(..., L) -> (..., M+1)``; kernel ``dsb200_acorr`` (time-domain lag sums, no FFT).

    The reference goes through ``rfft -> |.|^2 -> irfft`` (acorr.py:112-121); with ``M << L`` the M + 1 lag sums
    ``r_k = sum_n x_n x_{n+k}`` are cheaper computed directly: one warp per frame, lanes over lags, float64
    accumulators (speech frames are badly conditioned for the Levinson solve that follows), formatter applied in
    the same pass.  From the waveform the fused ``lpc_from_waveform`` keeps a 25-deep sliding register window
    instead (``csrc/fused_wave.cu``).  Backward: ``dsb200_acorr_backward``,
    ``gx_n = sum_k gr_k (x_{n+k} + x_{n-k})``.
    """

    _takes_input_size = True

    def __init__(self, frame_length: int, acr_order: int, out_format: str | int = "naive") -> None:
        super().__init__()
        self.in_dim = frame_length
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        check_size(x.size(-1), self.in_dim, "length of waveform")
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = Autocorrelation._precompute(x.size(-1), *args, **kwargs)
        return Autocorrelation._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(frame_length: int, acr_order: int) -> None:
        if frame_length <= 0:
            raise ValueError("frame_length must be positive.")
        if frame_length <= acr_order:
            raise ValueError("acr_order must be less than frame_length.")

    @staticmethod
    def _precompute(frame_length: int, acr_order: int, out_format: str | int = "naive") -> Precomputed:
        Autocorrelation._check(frame_length, acr_order)
        if isinstance(out_format, str) and out_format in _FORMATS:
            fmt = _FORMATS[out_format]
        elif isinstance(out_format, int) and not isinstance(out_format, bool) and 0 <= out_format <= 3:
            fmt = out_format
        else:
            raise ValueError(f"out_format {out_format} is not supported.")
        return Precomputed(values={"acr_order": acr_order, "out_format": fmt})

    @staticmethod
    def _forward(x: torch.Tensor, *, acr_order: int, out_format: int) -> torch.Tensor:
        return ops.acorr(x, acr_order, out_format)
