"""Mel filter-bank analysis (drop-in for diffsptk/modules/fbank.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import check_size, filter_values
from .base import BaseFunctionalModule, Precomputed

_FORMATS = {"y": 0, "yE": 1, "y,E": 2}


def fbank_check(fft_length, n_channel, sample_rate, f_min, f_max, floor, gamma, erb_factor) -> None:
    if fft_length <= 1:
        raise ValueError("fft_length must be greater than 1.")
    if n_channel <= 0:
        raise ValueError("n_channel must be positive.")
    if sample_rate <= 0:
        raise ValueError("sample_rate must be positive.")
    if f_min < 0 or sample_rate / 2 <= f_min:
        raise ValueError("Invalid f_min.")
    if f_max is not None and not (f_min < f_max <= sample_rate / 2):
        raise ValueError("Invalid f_min and f_max.")
    if floor <= 0:
        raise ValueError("floor must be positive.")
    if 1 < abs(gamma):
        raise ValueError("gamma must be in [-1, 1].")
    if erb_factor is not None and erb_factor <= 0:
        raise ValueError("erb_factor must be positive.")


def support_of(H: torch.Tensor, begin=None, end=None):
    """Column support to hand to the kernel: the precomputed one for a fixed filter bank, none (dense
    walk) for a trainable one, computed on the fly when the caller brought its own matrix."""
    if isinstance(H, torch.nn.Parameter):
        return None, None
    if begin is not None and end is not None:
        return begin, end
    return tables.column_support(H)


class MelFilterBankAnalysis(BaseFunctionalModule):
    """``(..., L/2+1) -> (..., C)`` [+ energy]; kernel ``dsb200_fbank``; buffer ``H`` as in the reference.

    The kernel walks only the non-zero rows of each triangular filter (493 of 10 280 weights at
    512 / 40 / 16 kHz) instead of the dense ``sqrt(x) @ H`` of fbank.py:315-316.
    """

    _takes_input_size = True

    def __init__(self, *, fft_length: int, n_channel: int, sample_rate: int, f_min: float = 0,
                 f_max: float | None = None, floor: float = 1e-5, gamma: float = 0, scale: str = "htk",
                 erb_factor: float | None = None, use_power: bool = False, out_format: str | int = "y",
                 learnable: bool = False, device: torch.device | None = None,
                 dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = fft_length // 2 + 1
        pre = self._precompute(**filter_values(locals(), drop_keys=["learnable"]))
        support = {k: pre.tensors.pop(k) for k in ("H_begin", "H_end")}
        self._register_precomputed(pre, learnable=learnable)
        if not learnable:  # a trainable H may grow new non-zeros: the kernel then walks it densely
            for k, v in support.items():
                self.register_buffer(k, v, persistent=False)

    def forward(self, x: torch.Tensor):
        check_size(x.size(-1), self.in_dim, "dimension of spectrum")
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs):
        pre = MelFilterBankAnalysis._precompute(2 * x.size(-1) - 2, *args, **kwargs, device=x.device,
                                                dtype=x.dtype)
        return MelFilterBankAnalysis._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(fft_length, n_channel, sample_rate, f_min, f_max, floor, gamma, erb_factor) -> None:
        fbank_check(fft_length, n_channel, sample_rate, f_min, f_max, floor, gamma, erb_factor)

    @staticmethod
    def _precompute(fft_length: int, n_channel: int, sample_rate: int, f_min: float, f_max: float | None,
                    floor: float, gamma: float, scale: str, erb_factor: float | None, use_power: bool,
                    out_format: str | int, device: torch.device | None,
                    dtype: torch.dtype | None) -> Precomputed:
        MelFilterBankAnalysis._check(fft_length, n_channel, sample_rate, f_min, f_max, floor, gamma, erb_factor)
        if isinstance(out_format, str) and out_format in _FORMATS:
            fmt = _FORMATS[out_format]
        elif isinstance(out_format, int) and not isinstance(out_format, bool) and 0 <= out_format <= 2:
            fmt = out_format
        else:
            raise ValueError(f"out_format {out_format} is not supported.")
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        H = tables.make_fbank_matrix(fft_length, n_channel, sample_rate, f_min, f_max, scale, erb_factor,
                                     device, dtype)
        begin, end = tables.column_support(H)
        return Precomputed(values=dict(floor=floor, gamma=gamma, use_power=use_power, out_format=fmt),
                           tensors={"H": H, "H_begin": begin, "H_end": end})

    @staticmethod
    def _forward(x: torch.Tensor, *, floor: float, gamma: float, use_power: bool, out_format: int,
                 H: torch.Tensor, H_begin: torch.Tensor | None = None, H_end: torch.Tensor | None = None):
        cb, ce = support_of(H, H_begin, H_end)
        y, E = ops.fbank(x, H, cb, ce, floor, gamma, use_power, out_format != 0)
        if out_format == 0:
            return y
        if out_format == 1:
            return torch.cat((y, E), dim=-1)
        return y, E
