"""Mel-generalized cepstral analysis (drop-in for diffsptk/modules/mgcep.py).

SURVEY.md section 8(f) rank 3.  ``gamma = 0`` is the fused mel-cepstral kernel (``dsb200_mcep``).  For
``-1 <= gamma < 0`` the reference's Newton iteration (mgcep.py:173-246) runs step by step on this package's kernels:
the warping matrices (``cfreqt`` / ``pfreqt`` / ``rfreqt`` / ``ptrans`` / ``qtrans``: ``dsb200_rowmat`` with the
reference's tables), the transforms (``dsb200_rfft``, ``dsb200_ifftr`` -- no torch.fft), the
Toeplitz-plus-Hankel solve (``dsb200_thsolve`` instead of torch.linalg.solve) and the coefficient conversions
(``dsb200_rowconv``, ``mgc2mgc``); the pointwise spectrum algebra between them is torch elementwise ops on the
device.  First version: one launch per step, like the reference; a fused per-frame kernel in the style of
``mcep_fast.cu`` is the natural next step.
"""

from __future__ import annotations

import torch
from torch import nn

from .. import ops, tables
from ..utils import check_size, get_gamma
from .b2mc import MLSADigitalFilterCoefficientsToMelCepstrum
from .gnorm import GeneralizedCepstrumGainNormalization
from .ignorm import GeneralizedCepstrumInverseGainNormalization
from .mc2b import MelCepstrumToMLSADigitalFilterCoefficients
from .mcep import MelCepstralAnalysis
from .mgc2mgc import MelGeneralizedCepstrumToMelGeneralizedCepstrum


class _RowMat(nn.Module):
    """``y = x @ A`` with a fixed table (buffer ``A``, non-persistent like the reference's)."""

    def __init__(self, A: torch.Tensor) -> None:
        super().__init__()
        self.register_buffer("A", A, persistent=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.rowmat(x, self.A)


class MelGeneralizedCepstralAnalysis(nn.Module):
    """``(..., L/2+1)`` power spectrum ``-> (..., M+1)`` mel-generalized cepstrum."""

    def __init__(self, *, fft_length: int, cep_order: int, alpha: float = 0, gamma: float = 0, c: int | None = None,
                 n_iter: int = 0, device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        gamma = get_gamma(gamma, c)
        if fft_length <= 1:
            raise ValueError("fft_length must be greater than 1.")
        if cep_order < 0:
            raise ValueError("cep_order must be non-negative.")
        if fft_length < 2 * cep_order:
            raise ValueError("cep_order must be less than or equal to fft_length // 2.")
        if 1 <= abs(alpha):
            raise ValueError("alpha must be in (-1, 1).")
        if gamma < -1 or 0 < gamma:
            raise ValueError("gamma must be in [-1, 0].")
        if n_iter < 0:
            raise ValueError("n_iter must be non-negative.")
        self.fft_length = fft_length
        self.cep_order = cep_order
        self.gamma = gamma
        self.n_iter = n_iter
        if gamma == 0:
            self.mcep = MelCepstralAnalysis(fft_length=fft_length, cep_order=cep_order, alpha=alpha, n_iter=n_iter,
                                            device=device, dtype=dtype)
            return
        if fft_length % 2:
            raise NotImplementedError("fft_length must be even (the kernels transform real sequences by the "
                                      "half-length trick).")
        M = cep_order
        self.cfreqt = _RowMat(tables.make_mgcep_freqt_matrix(M, fft_length - 1, -alpha, device, dtype))
        self.pfreqt = _RowMat(tables.make_mgcep_freqt_matrix(fft_length - 1, 2 * M, alpha, device, dtype))
        self.rfreqt = _RowMat(tables.make_mgcep_freqt_matrix(fft_length - 1, M, alpha, device, dtype))
        self.ptrans = _RowMat(tables.make_mgcep_ptrans(2 * M, alpha, device, dtype))
        self.qtrans = _RowMat(tables.make_mgcep_qtrans(2 * M, alpha, device, dtype))
        self.b2b = nn.Sequential(
            GeneralizedCepstrumInverseGainNormalization(M, -1),
            MLSADigitalFilterCoefficientsToMelCepstrum(M, alpha, device=device, dtype=dtype),
            MelGeneralizedCepstrumToMelGeneralizedCepstrum(M, M, in_gamma=-1, out_gamma=gamma, device=device,
                                                           dtype=dtype),
            MelCepstrumToMLSADigitalFilterCoefficients(M, alpha, device=device, dtype=dtype),
            GeneralizedCepstrumGainNormalization(M, gamma),
        )
        self.b2mc = nn.Sequential(
            GeneralizedCepstrumInverseGainNormalization(M, gamma),
            MLSADigitalFilterCoefficientsToMelCepstrum(M, alpha, device=device, dtype=dtype),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.gamma == 0:
            return self.mcep(x)
        M, L = self.cep_order, self.fft_length
        check_size(x.size(-1), L // 2 + 1, "dimension of spectrum")
        if not x.dtype.is_floating_point:
            x = x.to(torch.get_default_dtype())
        if M == 0:
            raise NotImplementedError("cep_order = 0 is not supported for gamma != 0.")

        def irfft(re, im=None):   # torch.fft.irfft of a half-spectrum -> L real samples
            return ops.ifftr(torch.complex(re, torch.zeros_like(re) if im is None else im), L)

        def newton(gamma, b1):
            b = torch.cat((torch.zeros_like(b1[..., :1]), b1), dim=-1)
            C = torch.view_as_complex(ops.rfft(self.cfreqt(b), L, 0))
            if gamma == -1:
                p_re = x
            else:
                X = 1 + gamma * C.real
                Y = gamma * C.imag
                XX, YY = X * X, Y * Y
                D = XX + YY
                p_re = x * torch.pow(D, -1 / gamma) / D
                q = p_re / D
                q_re, q_im = q * (XX - YY), q * (2 * X * Y)
                r_re, r_im = p_re * X, p_re * Y
            p = self.pfreqt(irfft(p_re))
            if gamma == -1:
                q = p
                r = p[..., : M + 1]
            else:
                q = self.pfreqt(irfft(q_re, q_im))
                r = self.rfreqt(irfft(r_re, r_im))
            p = self.ptrans(p)
            q = self.qtrans(q)
            if gamma != -1:
                eps = r[..., 0] + gamma * (r[..., 1:] * b1).sum(-1)
            # (symmetric_toeplitz(p[:M]) + hankel(q[2:] (1 + gamma))) gradient = r[1:]
            b1 = b1 + ops.thsolve(p[..., :M], q[..., 2:] * (1 + gamma), r[..., 1:])
            if gamma == -1:
                eps = r[..., 0] + gamma * (r[..., 1:] * b1).sum(-1)
            return torch.sqrt(eps).unsqueeze(-1), b1

        b1 = torch.zeros(*x.shape[:-1], M, device=x.device, dtype=x.dtype)
        b0, b1 = newton(-1, b1)
        if self.gamma != -1:
            b = self.b2b(torch.cat((b0, b1), dim=-1))
            b1 = b[..., 1:]
            for _ in range(self.n_iter):
                b0, b1 = newton(self.gamma, b1)
        return self.b2mc(torch.cat((b0, b1), dim=-1))
