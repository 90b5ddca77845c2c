"""Mel-cepstral analysis (drop-in for diffsptk/modules/mcep.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import check_size, filter_values, get_layer
from .base import BaseFunctionalModule, Precomputed
from .freqt import FrequencyTransform, _warp_check


class MelCepstralAnalysis(BaseFunctionalModule):
    """``(..., L/2+1) power spectrum -> (..., M+1)``; kernel ``dsb200_mcep``.

    The whole Newton iteration of mcep.py:209-222 stays on chip per frame.  The reference's
    ``ifreqt -> rfft.real`` and ``irfft -> rfreqt`` pairs are linear, so they are folded on the host
    (float64) into two dense tables ``G`` and ``Hm``; the initial ``irfft -> freqt`` likewise into
    ``P0`` (``tables.make_mcep_tables``).  The sub-layers ``freqt`` / ``ifreqt`` / ``rfreqt`` and the
    ``alpha_vector`` buffer are kept under the reference's names.
    """

    _takes_input_size = True

    def __init__(self, *, fft_length: int, cep_order: int, alpha: float = 0, n_iter: int = 0,
                 device: torch.device | None = None, dtype: torch.dtype | None = None):
        super().__init__()
        self.in_dim = fft_length // 2 + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor):
        check_size(x.size(-1), self.in_dim, "dimension of spectrum")
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs):
        pre = MelCepstralAnalysis._precompute(2 * x.size(-1) - 2, *args, **kwargs, dtype=x.dtype,
                                              device=x.device, module=False)
        return MelCepstralAnalysis._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(fft_length: int, cep_order: int, alpha: float, n_iter: int):
        if fft_length <= 1:
            raise ValueError("fft_length must be greater than 1.")
        if cep_order < 0:
            raise ValueError("cep_order must be non-negative.")
        if fft_length < 2 * cep_order:
            raise ValueError("cep_order must be less than or equal to fft_length // 2.")
        if 1 <= abs(alpha):
            raise ValueError("alpha must be in (-1, 1).")
        if n_iter < 0:
            raise ValueError("n_iter must be non-negative.")

    @staticmethod
    def _precompute(fft_length: int, cep_order: int, alpha: float, n_iter: int,
                    device: torch.device | None, dtype: torch.dtype | None, module: bool = True):
        MelCepstralAnalysis._check(fft_length, cep_order, alpha, n_iter)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        P0, G, Hm = tables.make_mcep_tables(fft_length, cep_order, alpha, device, dtype)
        alpha_vector = (-alpha) ** torch.arange(cep_order + 1, device=device, dtype=dtype)
        layers = {}
        if module:
            H = fft_length // 2
            layers = {
                "freqt": get_layer(True, FrequencyTransform, dict(in_order=H, out_order=cep_order, alpha=alpha,
                                                                  device=device, dtype=dtype)),
                "ifreqt": get_layer(True, FrequencyTransform, dict(in_order=cep_order, out_order=H, alpha=-alpha,
                                                                   device=device, dtype=dtype)),
                "rfreqt": get_layer(True, CoefficientsFrequencyTransform,
                                    dict(in_order=H, out_order=2 * cep_order, alpha=alpha, device=device,
                                         dtype=dtype)),
            }
        return Precomputed(values={"fft_length": fft_length, "n_iter": n_iter}, layers=layers,
                           tensors={"alpha_vector": alpha_vector, "P0": P0, "G": G, "Hm": Hm})

    @staticmethod
    def _forward(x: torch.Tensor, *, fft_length: int, n_iter: int, alpha_vector: torch.Tensor,
                 P0: torch.Tensor, G: torch.Tensor, Hm: torch.Tensor, freqt=None, ifreqt=None,
                 rfreqt=None) -> torch.Tensor:
        return ops.mcep(x, P0, G, Hm, alpha_vector, n_iter)


class CoefficientsFrequencyTransform(BaseFunctionalModule):
    """mcep-internal warping of correlation-like sequences (mcep.py:227-288); buffer ``A``."""

    _takes_input_size = True

    def __init__(self, in_order: int, out_order: int, alpha: float = 0, device: torch.device | None = None,
                 dtype: torch.dtype | None = None):
        super().__init__()
        self.in_dim = in_order + 1
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, c: torch.Tensor) -> torch.Tensor:
        check_size(c.size(-1), self.in_dim, "dimension of cepstrum")
        return self._call_forward(c)

    @staticmethod
    def _func(c: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = CoefficientsFrequencyTransform._precompute(c.size(-1) - 1, *args, **kwargs, device=c.device,
                                                         dtype=c.dtype)
        return CoefficientsFrequencyTransform._apply_precomputed(pre, c=c)

    @staticmethod
    def _check(in_order: int, out_order: int, alpha: float):
        _warp_check(in_order, out_order, alpha)

    @staticmethod
    def _precompute(in_order: int, out_order: int, alpha: float, device: torch.device | None,
                    dtype: torch.dtype | None):
        CoefficientsFrequencyTransform._check(in_order, out_order, alpha)
        return Precomputed(tensors={"A": tables.make_coef_freqt_matrix(in_order, out_order, alpha, device, dtype)})

    @staticmethod
    def _forward(c: torch.Tensor, *, A: torch.Tensor) -> torch.Tensor:
        return ops.rowmat(c, A)
