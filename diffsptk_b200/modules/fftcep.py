"""Cepstral analysis by the improved cepstral method (drop-in for diffsptk/modules/fftcep.py).

SURVEY.md section 8(f) rank 3, first version: the transforms (one inverse real FFT, then one Hermitian FFT pair
per iteration) run on this package's kernels -- ``dsb200_ifftr`` and ``dsb200_rfft``, no torch.fft -- while the
elementwise steps in between (log, clamp at zero, the accelerated update) are torch elementwise ops on the
device (so autograd flows through the composite: every transform has a native adjoint).  A single fused
kernel per iteration is the natural next step.
"""

from __future__ import annotations

import torch

from .. import ops
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed


class CepstralAnalysis(BaseFunctionalModule):
    """``(..., L/2+1)`` power spectrum ``-> (..., M+1)`` cepstrum (fftcep.py:116-136)."""

    _takes_input_size = True

    def __init__(self, *, fft_length: int, cep_order: int, accel: float = 0, n_iter: int = 0) -> None:
        super().__init__()
        self.in_dim = fft_length // 2 + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        check_size(x.size(-1), self.in_dim, "dimension of spectrum")
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = CepstralAnalysis._precompute(2 * x.size(-1) - 2, *args, **kwargs)
        return CepstralAnalysis._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(fft_length: int, cep_order: int, accel: float, n_iter: int) -> None:
        if fft_length <= 1:
            raise ValueError("fft_length must be greater than 1.")
        if cep_order < 0:
            raise ValueError("cep_order must be non-negative.")
        if fft_length < 2 * cep_order:
            raise ValueError("cep_order must be less than or equal to fft_length // 2.")
        if accel < 0:
            raise ValueError("accel must be non-negative.")
        if n_iter < 0:
            raise ValueError("n_iter must be non-negative.")

    @staticmethod
    def _precompute(fft_length: int, cep_order: int, accel: float, n_iter: int) -> Precomputed:
        CepstralAnalysis._check(fft_length, cep_order, accel, n_iter)
        return Precomputed(values={"cep_order": cep_order, "accel": accel, "n_iter": n_iter})

    @staticmethod
    def _forward(x: torch.Tensor, *, cep_order: int, accel: float, n_iter: int) -> torch.Tensor:
        if not x.dtype.is_floating_point:
            x = x.to(torch.get_default_dtype())
        n_bins = x.size(-1)
        n_fft, n_cep = 2 * (n_bins - 1), cep_order + 1
        as_half_spectrum = lambda t: torch.complex(t, torch.zeros_like(t))  # noqa: E731  (real, even sequence)

        # cepstrum of the log spectrum; its first M + 1 terms are the estimate, the tail is the residual
        cep = ops.ifftr(as_half_spectrum(torch.log(x)), n_fft)
        estimate = cep[..., :n_cep]
        residual = torch.cat((torch.zeros_like(cep[..., :n_cep]), cep[..., n_cep:n_bins]), -1)
        gain = 1 + accel
        for _ in range(n_iter):
            # residual -> log-spectral domain (hfft of a real sequence = n * irfft), keep only what lies above
            # the current envelope, and come back (ihfft(.).real = rfft(.).real / n)
            above = (ops.ifftr(as_half_spectrum(residual), n_fft) * n_fft).clamp(min=0)
            residual = ops.rfft(above, n_fft, 1) / n_fft
            step = residual[..., :n_cep] * gain
            estimate = estimate + step
            residual = torch.cat((residual[..., :n_cep] - step, residual[..., n_cep:]), -1)
        halve = torch.ones(n_cep, device=x.device, dtype=estimate.dtype)   # out of place: autograd flows through
        halve[0] = 0.5
        if n_bins == n_cep:
            halve[n_cep - 1] = 0.5
        return estimate * halve
