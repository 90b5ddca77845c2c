"""LPC -> line spectral pairs (drop-in for diffsptk/modules/lpc2lsp.py)."""

from __future__ import annotations

import math

import torch

from .. import ops
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed

TAU = 2 * math.pi


class LinearPredictiveCoefficientsToLineSpectralPairs(BaseFunctionalModule):
    """``(..., M+1) -> (..., M+1)``: ``[K, a_1..a_M] -> [K, w_1..w_M]``; kernel ``dsb200_lpc2lsp`` (csrc/lsp.cu).

    The reference takes the eigenvalues of two companion matrices (lpc2lsp.py:178-191).  The roots it is after
    all lie on the unit circle, so the kernel finds them there: a warp evaluates the two Chebyshev series on a
    grid of angles, brackets the sign changes and bisects each bracket in float64.  Differentiable by implicit
    differentiation of the two real functions whose zeros the line spectral frequencies are.
    """

    _takes_input_size = True

    def __init__(self, lpc_order: int, log_gain: bool = False, sample_rate: int | None = None,
                 out_format: str | int = "radian", device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = lpc_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, a: torch.Tensor) -> torch.Tensor:
        check_size(a.size(-1), self.in_dim, "dimension of LPC")
        return self._call_forward(a)

    @staticmethod
    def _func(a: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = LinearPredictiveCoefficientsToLineSpectralPairs._precompute(a.size(-1) - 1, *args, **kwargs,
                                                                          device=a.device, dtype=a.dtype)
        return LinearPredictiveCoefficientsToLineSpectralPairs._apply_precomputed(pre, a=a)

    @staticmethod
    def _check(lpc_order: int, log_gain: bool, sample_rate: int | None, out_format: str | int) -> None:
        if lpc_order < 0:
            raise ValueError("lpc_order must be non-negative.")
        if out_format in (2, 3, "hz", "khz") and (sample_rate is None or sample_rate <= 0):
            raise ValueError("sample_rate must be positive.")

    @staticmethod
    def _precompute(lpc_order: int, log_gain: bool, sample_rate: int | None, out_format: str | int,
                    device: torch.device | None, dtype: torch.dtype | None) -> Precomputed:
        LinearPredictiveCoefficientsToLineSpectralPairs._check(lpc_order, log_gain, sample_rate, out_format)
        if out_format in (0, "radian"):
            scale = 1.0
        elif out_format in (1, "cycle"):
            scale = 1 / TAU
        elif out_format in (2, "khz"):
            scale = 1 / (TAU / sample_rate * 1000)
        elif out_format in (3, "hz"):
            scale = 1 / (TAU / sample_rate)
        else:
            raise ValueError(f"out_format {out_format} is not supported.")
        return Precomputed(values={"log_gain": log_gain, "scale": scale})

    @staticmethod
    def _forward(a: torch.Tensor, *, log_gain: bool, scale: float) -> torch.Tensor:
        return ops.lpc2lsp(a, log_gain, scale)
