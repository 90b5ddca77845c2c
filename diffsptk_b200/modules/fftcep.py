"""Cepstral analysis by the improved cepstral method (drop-in for diffsptk/modules/fftcep.py).

SURVEY.md section 8(f) rank 3, first version: the transforms (one inverse real FFT, then one Hermitian FFT pair
per iteration) run on this package's kernels -- ``dsb200_ifftr`` and ``dsb200_rfft``, no torch.fft -- while the
elementwise steps in between (log, clamp at zero, the accelerated update) are torch elementwise ops on the
device.  A single fused kernel per iteration is the natural next step.
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
        ops._no_grad_check(x)
        N, H = cep_order + 1, x.size(-1)
        n = 2 * (H - 1)
        if not x.dtype.is_floating_point:
            x = x.to(torch.get_default_dtype())

        def hermitian(t):  # real, even sequence of H points as the half spectrum the inverse kernel takes
            return torch.complex(t, torch.zeros_like(t))

        e_full = ops.ifftr(hermitian(torch.log(x)), n)                  # irfft(log x)
        v = e_full[..., :N].clone()
        e = torch.zeros_like(x)
        e[..., N:H] = e_full[..., N:H]
        for _ in range(n_iter):
            E = ops.ifftr(hermitian(e), n) * n                           # hfft(e) of a real sequence
            E.clamp_(min=0)
            e = ops.rfft(E, n, 1) / n                                    # ihfft(E).real
            t = e[..., :N] * (1 + accel)
            v += t
            e[..., :N] -= t
        v[..., 0] *= 0.5
        if H == N:
            v[..., N - 1] *= 0.5
        return v
