"""Mel-generalized cepstrum -> spectrum (drop-in for diffsptk/modules/mgc2sp.py)."""

from __future__ import annotations

import math
from typing import Callable

import torch

from .. import ops
from ..utils import check_size, filter_values, get_layer
from .base import BaseFunctionalModule, Precomputed
from .mgc2mgc import MelGeneralizedCepstrumToMelGeneralizedCepstrum


_FORMAT_IDS = {"db": 0, "log-magnitude": 1, "magnitude": 2, "power": 3, "cycle": 4, "radian": 5, "degree": 6}


def _log_spectrum_formatter(out_format) -> Callable:
    """What to return from the complex log spectrum ``log|H| + j arg H`` (mgc2sp.py:152-169)."""
    if out_format == "complex":
        return lambda sp: torch.polar(torch.exp(sp.real), sp.imag)
    fid = _FORMAT_IDS.get(out_format, out_format) if isinstance(out_format, str) else out_format
    if isinstance(fid, bool) or fid not in range(7):
        raise ValueError(f"out_format {out_format} is not supported.")
    amplitude = {0: 20 / math.log(10), 1: 1.0}                        # scaled log magnitude
    phase = {4: 1 / math.pi, 5: 1.0, 6: 180 / math.pi}                # scaled phase
    if fid in amplitude:
        return lambda sp: sp.real * amplitude[fid] if fid == 0 else sp.real
    if fid in phase:
        return lambda sp: sp.imag / torch.pi if fid == 4 else (sp.imag if fid == 5 else sp.imag * (180 / torch.pi))
    return lambda sp: torch.exp(sp.real if fid == 2 else 2 * sp.real)


class MelGeneralizedCepstrumToSpectrum(BaseFunctionalModule):
    """``(..., M+1) -> (..., L/2+1)``: un-warp / un-gamma the cepstrum to a plain cepstrum of order ``L/2``
    (``mgc2mgc``), transform it with ``dsb200_rfft`` and format the log spectrum (mgc2sp.py:131-202)."""

    _takes_input_size = True

    def __init__(self, cep_order: int, fft_length: int, *, alpha: float = 0, gamma: float = 0, norm: bool = False,
                 mul: bool = False, n_fft=512, out_format: str | int = "power",
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self.in_dim = cep_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, mc: torch.Tensor) -> torch.Tensor:
        check_size(mc.size(-1), self.in_dim, "dimension of cepstrum")
        return self._call_forward(mc)

    @staticmethod
    def _func(mc: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = MelGeneralizedCepstrumToSpectrum._precompute(mc.size(-1) - 1, *args, **kwargs, dtype=mc.dtype,
                                                           device=mc.device, module=False)
        return MelGeneralizedCepstrumToSpectrum._apply_precomputed(pre, mc=mc)

    @staticmethod
    def _check() -> None:
        pass

    @staticmethod
    def _precompute(cep_order: int, fft_length: int, alpha: float, gamma: float, norm: bool, mul: bool, n_fft: int,
                    out_format: str | int, device: torch.device | None, dtype: torch.dtype | None,
                    module: bool = True) -> Precomputed:
        MelGeneralizedCepstrumToSpectrum._check()
        formatter = _log_spectrum_formatter(out_format)
        mgc2c = get_layer(module, MelGeneralizedCepstrumToMelGeneralizedCepstrum,
                          dict(in_order=cep_order, in_alpha=alpha, in_gamma=gamma, in_norm=norm, in_mul=mul,
                               out_order=fft_length // 2, out_alpha=0, out_gamma=0, out_norm=False, out_mul=False,
                               n_fft=n_fft, device=device, dtype=dtype))
        return Precomputed(values={"formatter": formatter}, layers={"mgc2c": mgc2c})

    @staticmethod
    def _forward(mc: torch.Tensor, *, formatter: Callable, mgc2c: Callable) -> torch.Tensor:
        c = mgc2c(mc)
        sp = torch.view_as_complex(ops.rfft(c, (c.size(-1) - 1) * 2, 0))
        return formatter(sp)
