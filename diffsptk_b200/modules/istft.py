"""Inverse short-time Fourier transform as ONE fused kernel (drop-in for diffsptk/modules/istft.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import filter_values, get_layer
from .base import BaseFunctionalModule, Precomputed
from .ifftr import RealValuedInverseFastFourierTransform
from .stft import LEARNABLES, ShortTimeFourierTransform
from .unframe import Unframe


class InverseShortTimeFourierTransform(BaseFunctionalModule):
    """complex ``(..., T/P, N/2+1) -> (..., T)``.

    The sub-layers ``ifftr`` / ``unframe`` exist under the reference's names (``istft.unframe.window``), but
    ``forward`` launches ``dsb200_istft``: inverse FFT, synthesis window, overlap-add and the sum-of-squares
    normalisation in one pass, without the ``[..., N, L]`` frame tensor in HBM.
    """

    def __init__(self, frame_length: int, frame_period: int, fft_length: int, *, center: bool = True,
                 window: str = "blackman", norm: str = "power", symmetric: bool = True,
                 learnable: bool | list[str] = False, device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, y: torch.Tensor, out_length: int | None = None) -> torch.Tensor:
        return self._call_forward(y, out_length)

    @staticmethod
    def _func(y: torch.Tensor, out_length: int | None, *args, **kwargs) -> torch.Tensor:
        dt = torch.float64 if y.dtype == torch.complex128 else torch.float32
        pre = InverseShortTimeFourierTransform._precompute(*args, **kwargs, learnable=False, device=y.device,
                                                           dtype=dt, module=False)
        return InverseShortTimeFourierTransform._apply_precomputed(pre, y=y, out_length=out_length)

    @staticmethod
    def _check(*args, **kwargs) -> None:
        ShortTimeFourierTransform._check(*args, **kwargs)

    @staticmethod
    def _precompute(frame_length: int, frame_period: int, fft_length: int, center: bool, window: str, norm: str,
                    symmetric: bool, learnable: bool | list[str], device: torch.device | None,
                    dtype: torch.dtype | None, module: bool = True) -> Precomputed:
        InverseShortTimeFourierTransform._check(learnable)
        keys = LEARNABLES if learnable is True else (() if learnable is False else tuple(learnable))
        basis = "basis" in keys and module
        RealValuedInverseFastFourierTransform._check(fft_length, frame_length)
        values = dict(frame_period=frame_period, center=center, fft_length=fft_length, basis=basis)
        un_params = dict(frame_length=frame_length, frame_period=frame_period, center=center, window=window,
                         norm=norm, symmetric=symmetric, learnable="window" in keys, device=device, dtype=dtype)
        if not module:
            un_params.pop("learnable")
            table = Unframe._precompute(**un_params).tensors["window"]
            return Precomputed(values=values, tensors={"window_table": table})
        ifftr = get_layer(True, RealValuedInverseFastFourierTransform,
                          dict(fft_length=fft_length, out_length=frame_length, learnable=basis, device=device,
                               dtype=dtype))
        unframe = get_layer(True, Unframe, un_params)
        return Precomputed(values=values, layers={"ifftr": ifftr, "unframe": unframe})

    @staticmethod
    def _forward(y: torch.Tensor, out_length: int | None, *, frame_period: int, center: bool, fft_length: int,
                 basis: bool = False, ifftr=None, unframe=None,
                 window_table: torch.Tensor | None = None) -> torch.Tensor:
        if basis:   # trainable inverse DFT basis: the reference's cascade (istft.py:186-193)
            return unframe(ifftr(y), out_length=out_length)
        table = (window_table if window_table is not None else unframe.window).reshape(-1)
        if not y.is_complex():
            raise ValueError("the input spectrogram must be complex")
        if y.dim() <= 1:
            raise ValueError("Input must be at least 2D tensor.")
        if 2 * (y.size(-1) - 1) != fft_length:
            raise ValueError(f"Unexpected dimension of spectrum (input {y.size(-1)} vs target {fft_length // 2 + 1}).")
        T = ops.unframe_length(y.size(-2), table.shape[-1], frame_period, center, out_length)
        return ops.istft(y, table, T, frame_period, center)
