"""Inverse real FFT (drop-in for diffsptk/modules/ifftr.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed


class RealValuedInverseFastFourierTransform(BaseFunctionalModule):
    """complex ``(..., L/2+1) -> (..., N)``; kernel ``dsb200_ifftr`` (no cuFFT / torch.fft on the path).

    One warp per spectrum row.  For power-of-two lengths the row is folded into the half-length complex
    sequence ``E + iO`` (``E``/``O`` = spectra of the even / odd samples), transformed by the radix-4 Stockham
    passes of ``csrc/rowfft.cuh`` as ``conj(FFT(conj(.)))`` and de-interleaved; other even lengths use a
    table-driven direct sum.  Like ``torch.fft.irfft`` the imaginary parts of the DC and Nyquist bins are
    ignored.  Only the first ``out_length`` samples are produced.  The fused inverse STFT never calls this op:
    it keeps the frames on chip (``csrc/istft512.cu``, ``csrc/inverse.cu``).
    """

    _takes_input_size = True

    def __init__(self, fft_length: int, out_length: int | None = None, learnable: bool = False,
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = fft_length // 2 + 1
        self._register_precomputed(self._precompute(**filter_values(locals())), learnable=learnable is True)

    def forward(self, y: torch.Tensor) -> torch.Tensor:
        check_size(y.size(-1), self.in_dim, "length of spectrum")
        return self._call_forward(y)

    @staticmethod
    def _func(y: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = RealValuedInverseFastFourierTransform._precompute(2 * y.size(-1) - 2, *args, **kwargs,
                                                                learnable=False, device=y.device, dtype=None)
        return RealValuedInverseFastFourierTransform._apply_precomputed(pre, y=y)

    @staticmethod
    def _check(fft_length: int, out_length: int | None) -> None:
        if fft_length <= 0 or fft_length % 2 == 1:
            raise ValueError("fft_length must be positive even.")
        if out_length is not None and (out_length <= 0 or fft_length < out_length):
            raise ValueError("out_length must be in [1, fft_length].")

    @staticmethod
    def _precompute(fft_length: int, out_length: int | None, learnable: bool, device: torch.device | None,
                    dtype: torch.dtype | None) -> Precomputed:
        RealValuedInverseFastFourierTransform._check(fft_length, out_length)
        tensors = {}
        if learnable:
            # the reference switches to a trainable inverse-DFT matrix (ifftr.py:117-124): [2 (L/2+1), out_length]
            W = torch.fft.ifft(torch.eye(fft_length, dtype=torch.double))[: fft_length // 2 + 1, :out_length]
            W[1:-1] *= 2
            W = torch.cat([W.real, -W.imag], dim=0)
            tensors["W"] = W.to(device=device, dtype=dtype if dtype is not None and dtype.is_floating_point
                                else torch.get_default_dtype())
        return Precomputed(values={"fft_length": fft_length, "out_length": out_length}, tensors=tensors)

    @staticmethod
    def _forward(y: torch.Tensor, *, fft_length: int, out_length: int | None,
                 W: torch.Tensor | None = None) -> torch.Tensor:
        if not y.is_complex():
            raise ValueError("the input spectrum must be complex")
        if W is not None:   # trainable inverse basis: dense product on the native row-times-matrix kernel
            return ops.rowmat(torch.cat([y.real, y.imag], dim=-1).to(W.dtype), W)
        return ops.ifftr(y, fft_length if out_length is None else out_length)
