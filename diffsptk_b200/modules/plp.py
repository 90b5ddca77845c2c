"""Perceptual linear predictive coefficients analysis (drop-in for diffsptk/modules/plp.py).

SURVEY.md section 8(f) rank 3: a direct consumer of the STFT power spectrum that reuses the path's kernels --
filter bank (``dsb200_fbank``), Levinson-Durbin (``dsb200_levdur``), the LPC -> cepstrum conversion
(``mgc2mgc``: ``dsb200_rowconv`` + ``dsb200_rfft`` / ``dsb200_ifftr``) and the inverse real FFT of the
equal-loudness-weighted, cube-root-compressed filter-bank outputs (``dsb200_ifftr``, which is what
``torch.fft.hfft(..., norm="forward")`` of a real sequence computes).  The three elementwise steps between the
kernels are torch ops on the device.
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import torch

from .. import ops, tables
from ..utils import filter_values, get_layer
from .base import BaseFunctionalModule, Precomputed
from .fbank import MelFilterBankAnalysis
from .levdur import LevinsonDurbin
from .mgc2mgc import MelGeneralizedCepstrumToMelGeneralizedCepstrum


def _equal_loudness(n_channel, sample_rate, f_min, f_max, scale) -> np.ndarray:
    """Equal-loudness weight at the centre frequency of every filter-bank channel (plp.py:279-285), float64."""
    lo = tables._to_auditory(np.asarray(f_min), scale)
    hi = tables._to_auditory(np.asarray(sample_rate / 2 if f_max is None else f_max), scale)
    centres = (hi - lo) / (n_channel + 1) * np.arange(1, n_channel + 2) + lo
    f = tables._from_auditory(centres, scale)[:-1] ** 2
    return (f / (f + 1.6e5)) ** 2 * (f + 1.44e6) / (f + 9.61e6)


def _plp_lifter(plp_order, lifter, device) -> torch.Tensor:
    """1 + (L/2) sin(pi m / L) with the zeroth term fixed at 2 (plp.py:287-289), float64."""
    m = torch.arange(plp_order + 1, device=device, dtype=torch.double)
    v = 1 + (lifter / 2) * torch.sin((torch.pi / lifter) * m)
    v[0] = 2
    return v


_LAYOUTS = {0: "y", 1: "yE", 2: "yc", 3: "ycE"}


def _packer(out_format) -> Callable:
    """``y`` = coefficients without c0, ``c`` = c0, ``E`` = energy, concatenated in the order the name spells."""
    name = _LAYOUTS.get(out_format, out_format)
    if isinstance(out_format, bool) or name not in _LAYOUTS.values():
        raise ValueError(f"out_format {out_format} is not supported.")

    def pack(y, c, E):
        parts = {"y": y, "c": c, "E": E}
        return y if name == "y" else torch.cat([parts[ch] for ch in name], dim=-1)
    return pack


class PerceptualLinearPredictiveCoefficientsAnalysis(BaseFunctionalModule):
    """``(..., L/2+1)`` power spectrum ``-> (..., M [+1] [+1])`` (plp.py:303-320)."""

    _takes_input_size = True

    def __init__(self, *, fft_length: int, plp_order: int, n_channel: int, sample_rate: int,
                 compression_factor: float = 0.33, lifter: int = 1, f_min: float = 0, f_max: float | None = None,
                 floor: float = 1e-5, gamma: float = 0, scale: str = "htk", erb_factor: float | None = None,
                 n_fft: int = 512, out_format: str | int = "y", learnable: bool = False,
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = PerceptualLinearPredictiveCoefficientsAnalysis._precompute(
            2 * x.size(-1) - 2, *args, **kwargs, learnable=False, device=x.device, dtype=x.dtype, module=False)
        return PerceptualLinearPredictiveCoefficientsAnalysis._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(plp_order: int, n_channel: int, compression_factor: float, lifter: int) -> None:
        if plp_order < 0:
            raise ValueError("plp_order must be non-negative.")
        if n_channel <= plp_order:
            raise ValueError("plp_order must be less than n_channel.")
        if compression_factor <= 0:
            raise ValueError("compression_factor must be positive.")
        if lifter < 0:
            raise ValueError("lifter must be non-negative.")

    @staticmethod
    def _precompute(fft_length: int, plp_order: int, n_channel: int, sample_rate: int, compression_factor: float,
                    lifter: int, f_min: float, f_max: float | None, floor: float, gamma: float, scale: str,
                    erb_factor: float | None, n_fft: int, out_format: str | int, learnable: bool,
                    device: torch.device | None, dtype: torch.dtype | None, module: bool = True) -> Precomputed:
        PerceptualLinearPredictiveCoefficientsAnalysis._check(plp_order, n_channel, compression_factor, lifter)
        formatter = _packer(out_format)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        fbank = get_layer(module, MelFilterBankAnalysis,
                          dict(fft_length=fft_length, n_channel=n_channel, sample_rate=sample_rate, f_min=f_min,
                               f_max=f_max, floor=floor, gamma=gamma, scale=scale, erb_factor=erb_factor,
                               use_power=True, out_format="y,E", learnable=learnable, device=device, dtype=dtype))
        levdur = get_layer(module, LevinsonDurbin, dict(lpc_order=plp_order, eps=0, device=device, dtype=dtype))
        lpc2c = get_layer(module, MelGeneralizedCepstrumToMelGeneralizedCepstrum,
                          dict(in_order=plp_order, in_alpha=0, in_gamma=-1, in_norm=True, in_mul=True,
                               out_order=plp_order, out_alpha=0, out_gamma=0, out_norm=False, out_mul=False,
                               n_fft=n_fft, device=device, dtype=dtype))
        tensors = {"equal_loudness_curve": tables._cast(_equal_loudness(n_channel, sample_rate, f_min, f_max, scale),
                                                        device, dtype),
                   "liftering_vector": tables._cast(_plp_lifter(plp_order, lifter, device), None, dtype)}
        return Precomputed(values={"compression_factor": compression_factor, "formatter": formatter},
                           layers={"fbank": fbank, "levdur": levdur, "lpc2c": lpc2c}, tensors=tensors)

    @staticmethod
    def _forward(x: torch.Tensor, *, compression_factor: float, formatter: Callable, fbank: Callable,
                 levdur: Callable, lpc2c: Callable, equal_loudness_curve: torch.Tensor,
                 liftering_vector: torch.Tensor) -> torch.Tensor:
        y, E = fbank(x)
        y = (torch.exp(y) * equal_loudness_curve.to(y.device)) ** compression_factor
        y = torch.cat((y[..., :1], y, y[..., -1:]), dim=-1)                      # replicate1
        n = 2 * (y.size(-1) - 1)
        # hfft(y, norm="forward") of a real half-spectrum = irfft(y): the autocorrelation-like sequence
        r = ops.ifftr(torch.complex(y, torch.zeros_like(y)), n)[..., : liftering_vector.numel()]
        y = lpc2c(levdur(r)) * liftering_vector.to(r.device)
        c, y = torch.split(y, [1, y.size(-1) - 1], dim=-1)
        return formatter(y, c, E)
