"""MFCC analysis (drop-in for diffsptk/modules/mfcc.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import filter_values, get_layer
from .base import BaseFunctionalModule, Precomputed
from .dct import DiscreteCosineTransform
from .fbank import MelFilterBankAnalysis, fbank_check, support_of

_FORMATS = {"y": 0, "yE": 1, "yc": 2, "ycE": 3}


def mfcc_format_id(out_format) -> int:
    if isinstance(out_format, str) and out_format in _FORMATS:
        return _FORMATS[out_format]
    if isinstance(out_format, int) and not isinstance(out_format, bool) and 0 <= out_format <= 3:
        return out_format
    raise ValueError(f"out_format {out_format} is not supported.")


class MelFrequencyCepstralCoefficientsAnalysis(BaseFunctionalModule):
    """``(..., L/2+1) -> (..., M [+1] [+1])``; kernel ``dsb200_mfcc`` (filter bank, log, DCT-II,
    lifter and output packing in one pass).  Sub-layers ``fbank`` / ``dct`` and the
    ``liftering_vector`` buffer keep the reference's names (mfcc.py:199-241)."""

    _takes_input_size = True

    def __init__(self, *, fft_length: int, mfcc_order: int, n_channel: int, sample_rate: int, lifter: int = 1,
                 f_min: float = 0, f_max: float | None = None, floor: float = 1e-5, gamma: float = 0,
                 scale: str = "htk", erb_factor: float | None = None, out_format: str | int = "y",
                 learnable: bool = False, device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = fft_length // 2 + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = MelFrequencyCepstralCoefficientsAnalysis._precompute(
            2 * x.size(-1) - 2, *args, **kwargs, learnable=False, device=x.device, dtype=x.dtype, module=False)
        return MelFrequencyCepstralCoefficientsAnalysis._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(mfcc_order: int, n_channel: int, lifter: int) -> None:
        if mfcc_order < 0:
            raise ValueError("mfcc_order must be non-negative.")
        if n_channel <= mfcc_order:
            raise ValueError("mfcc_order must be less than n_channel.")
        if lifter < 0:
            raise ValueError("lifter must be non-negative.")

    @staticmethod
    def _precompute(fft_length: int, mfcc_order: int, n_channel: int, sample_rate: int, lifter: int,
                    f_min: float, f_max: float | None, floor: float, gamma: float, scale: str,
                    erb_factor: float | None, out_format: str | int, learnable: bool,
                    device: torch.device | None, dtype: torch.dtype | None, module: bool = True) -> Precomputed:
        MelFrequencyCepstralCoefficientsAnalysis._check(mfcc_order, n_channel, lifter)
        fbank_check(fft_length, n_channel, sample_rate, f_min, f_max, floor, gamma, erb_factor)
        fmt = mfcc_format_id(out_format)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        values = dict(floor=floor, gamma=gamma, out_format=fmt)
        liftering_vector = tables.make_lifter(mfcc_order, lifter, device, dtype)
        if not module:
            H = tables.make_fbank_matrix(fft_length, n_channel, sample_rate, f_min, f_max, scale, erb_factor,
                                         device, dtype)
            W = tables.make_dct_matrix(n_channel, 2, device, dtype)
            begin, end = tables.column_support(H)
            return Precomputed(values=values, tensors={"liftering_vector": liftering_vector, "H_table": H,
                                                       "W_table": W, "H_begin": begin, "H_end": end})
        fbank = get_layer(True, MelFilterBankAnalysis,
                          dict(fft_length=fft_length, n_channel=n_channel, sample_rate=sample_rate, f_min=f_min,
                               f_max=f_max, floor=floor, gamma=gamma, scale=scale, erb_factor=erb_factor,
                               use_power=False, out_format="y,E", learnable=learnable, device=device,
                               dtype=dtype))
        dct = get_layer(True, DiscreteCosineTransform, dict(dct_length=n_channel, dct_type=2, device=device,
                                                            dtype=dtype))
        return Precomputed(values=values, layers={"fbank": fbank, "dct": dct},
                           tensors={"liftering_vector": liftering_vector})

    @staticmethod
    def _forward(x: torch.Tensor, *, floor: float, gamma: float, out_format: int,
                 liftering_vector: torch.Tensor, fbank=None, dct=None, H_table: torch.Tensor | None = None,
                 W_table: torch.Tensor | None = None, H_begin: torch.Tensor | None = None,
                 H_end: torch.Tensor | None = None) -> torch.Tensor:
        H = H_table if H_table is not None else fbank.H
        if H_table is None:
            H_begin, H_end = getattr(fbank, "H_begin", None), getattr(fbank, "H_end", None)
        W = W_table if W_table is not None else dct.W
        if x.size(-1) != H.size(0):
            raise ValueError(f"Unexpected dimension of spectrum (input {x.size(-1)} vs target {H.size(0)}).")
        cb, ce = support_of(H, H_begin, H_end)
        return ops.mfcc(x, H, cb, ce, W, liftering_vector, floor, gamma, out_format)
