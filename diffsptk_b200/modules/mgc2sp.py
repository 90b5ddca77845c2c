"""Mel-generalized cepstrum -> spectrum (drop-in for diffsptk/modules/mgc2sp.py)."""

from __future__ import annotations

import math
from typing import Callable

import torch

from .. import ops
from ..utils import check_size, filter_values, get_layer
from .base import BaseFunctionalModule, Precomputed
from .mgc2mgc import MelGeneralizedCepstrumToMelGeneralizedCepstrum


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
        if out_format in (0, "db"):
            formatter = lambda x: x.real * (20 / math.log(10))  # noqa: E731
        elif out_format in (1, "log-magnitude"):
            formatter = lambda x: x.real  # noqa: E731
        elif out_format in (2, "magnitude"):
            formatter = lambda x: torch.exp(x.real)  # noqa: E731
        elif out_format in (3, "power"):
            formatter = lambda x: torch.exp(2 * x.real)  # noqa: E731
        elif out_format in (4, "cycle"):
            formatter = lambda x: x.imag / torch.pi  # noqa: E731
        elif out_format in (5, "radian"):
            formatter = lambda x: x.imag  # noqa: E731
        elif out_format in (6, "degree"):
            formatter = lambda x: x.imag * (180 / torch.pi)  # noqa: E731
        elif out_format == "complex":
            formatter = lambda x: torch.polar(torch.exp(x.real), x.imag)  # noqa: E731
        else:
            raise ValueError(f"out_format {out_format} is not supported.")
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
