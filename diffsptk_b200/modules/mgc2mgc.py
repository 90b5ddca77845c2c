"""Mel-generalized cepstrum -> mel-generalized cepstrum (drop-in for diffsptk/modules/mgc2mgc.py).

SURVEY.md section 8(f) rank 3/4.  The conversion is the reference's sequence of steps (mgc2mgc.py:209-297), each
step running on this package's kernels: frequency transform = ``dsb200_rowmat`` with the reference's table,
gain (de)normalisation = ``dsb200_rowconv``, and the gamma conversion (``GeneralizedCepstrumToGeneralizedCepstrum``,
mgc2mgc.py:327-364) = ``dsb200_gc2gc``: forward transform, pointwise complex power / log and inverse transform in
one kernel (no torch.fft, any ``n_fft``).  The one-coefficient scalings (gamma multiplication / division) are torch
elementwise ops on the device; autograd flows through the whole chain (the backward of ``gc2gc`` recomputes the
step on the differentiable FFT kernels).
"""

from __future__ import annotations

from typing import Callable

import torch

from .. import ops, tables
from ..utils import filter_values
from .base import BaseFunctionalModule, Precomputed


def gc2gc(c1: torch.Tensor, out_order: int, in_gamma: float, out_gamma: float, n_fft: int) -> torch.Tensor:
    """Generalized cepstrum (normalised) with ``in_gamma`` -> ``out_gamma`` in the spectral domain
    (mgc2mgc.py:327-364): one kernel, ``dsb200_gc2gc`` (csrc/gc2gc.cu)."""
    return ops.gc2gc(c1, out_order, in_gamma, out_gamma, n_fft)


def _scale_tail(g: float) -> Callable:          # GammaDivision / GammaMultiplication (mgc2mgc.py:367-411)
    def f(c):
        return torch.cat((c[..., :1], c[..., 1:] * g), dim=-1)
    return f


def _tail_div(gamma: float) -> Callable:
    def f(c):
        return torch.cat((c[..., :1], c[..., 1:] / gamma), dim=-1)
    return f


def _zeroth_div(gamma: float) -> Callable:      # ZerothGammaDivision (mgc2mgc.py:414-427)
    def f(c):
        return torch.cat(((c[..., :1] - 1) / gamma, c[..., 1:]), dim=-1)
    return f


def _zeroth_mul(gamma: float) -> Callable:      # ZerothGammaMultiplication (mgc2mgc.py:430-439)
    def f(c):
        return torch.cat((c[..., :1] * gamma + 1, c[..., 1:]), dim=-1)
    return f


class MelGeneralizedCepstrumToMelGeneralizedCepstrum(BaseFunctionalModule):
    """``(..., M1+1) -> (..., M2+1)``."""

    _takes_input_size = True

    def __init__(self, in_order: int, out_order: int, in_alpha: float = 0, out_alpha: float = 0,
                 in_gamma: float = 0, out_gamma: float = 0, in_norm: bool = False, out_norm: bool = False,
                 in_mul: bool = False, out_mul: bool = False, n_fft: int = 512,
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, mc: torch.Tensor) -> torch.Tensor:
        return self._call_forward(mc)

    @staticmethod
    def _func(mc: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = MelGeneralizedCepstrumToMelGeneralizedCepstrum._precompute(
            mc.size(-1) - 1, *args, **kwargs, device=mc.device, dtype=mc.dtype, module=False)
        return MelGeneralizedCepstrumToMelGeneralizedCepstrum._apply_precomputed(pre, mc=mc)

    @staticmethod
    def _check(in_order: int, out_order: int, in_alpha: float, out_alpha: float, in_gamma: float,
               out_gamma: float, in_mul: bool, n_fft: int) -> None:
        if in_order < 0:
            raise ValueError("in_order must be non-negative.")
        if out_order < 0:
            raise ValueError("out_order must be non-negative.")
        if 1 <= abs(in_alpha):
            raise ValueError("in_alpha must be in (-1, 1).")
        if 1 <= abs(out_alpha):
            raise ValueError("out_alpha must be in (-1, 1).")
        if 1 < abs(in_gamma):
            raise ValueError("in_gamma must be in [-1, 1].")
        if 1 < abs(out_gamma):
            raise ValueError("out_gamma must be in [-1, 1].")
        if n_fft <= max(in_order, out_order) + 1:
            raise ValueError("n_fft must be much larger than order of cepstrum.")
        if 0 == in_gamma and in_mul:
            raise ValueError("Invalid combination of in_gamma and in_mul.")

    @staticmethod
    def _precompute(in_order: int, out_order: int, in_alpha: float, out_alpha: float, in_gamma: float,
                    out_gamma: float, in_norm: bool, out_norm: bool, in_mul: bool, out_mul: bool, n_fft: int,
                    device: torch.device | None, dtype: torch.dtype | None, module: bool = True) -> Precomputed:
        MelGeneralizedCepstrumToMelGeneralizedCepstrum._check(in_order, out_order, in_alpha, out_alpha, in_gamma,
                                                              out_gamma, in_mul, n_fft)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        gnorm = lambda g: (lambda c: ops.rowconv(c, ops.CONV_GNORM, g))      # noqa: E731
        ignorm = lambda g: (lambda c: ops.rowconv(c, ops.CONV_IGNORM, g))    # noqa: E731
        convert = lambda o: (lambda c: gc2gc(c, o, in_gamma, out_gamma, n_fft))  # noqa: E731

        seq: list[Callable] = []
        if not in_norm and in_mul:
            seq.append(_zeroth_div(in_gamma))
        alpha = (out_alpha - in_alpha) / (1 - in_alpha * out_alpha)
        if 0 == alpha:
            if in_order == out_order and in_gamma == out_gamma:
                if not in_mul and out_mul:
                    seq.append(_scale_tail(in_gamma))
                if not in_norm and out_norm:
                    seq.append(gnorm(in_gamma))
                if in_norm and not out_norm:
                    seq.append(ignorm(out_gamma))
                if in_mul and not out_mul:
                    seq.append(_tail_div(out_gamma))
            else:
                if in_mul:
                    seq.append(_tail_div(in_gamma))
                if not in_norm:
                    seq.append(gnorm(in_gamma))
                seq.append(convert(out_order))
                if not out_norm:
                    seq.append(ignorm(out_gamma))
                if out_mul:
                    seq.append(_scale_tail(out_gamma))
        else:
            if in_mul:
                seq.append(_tail_div(in_gamma))
            if in_norm:
                seq.append(ignorm(in_gamma))
            A = {None: tables.make_freqt_matrix(in_order, out_order, alpha, device, dtype)}

            def warp(c):   # the table follows the input's device (it is not a registered buffer)
                if c.device not in A:
                    A[c.device] = A[None].to(c.device)
                return ops.rowmat(c, A[c.device])
            seq.append(warp)
            if out_norm or in_gamma != out_gamma:
                seq.append(gnorm(in_gamma))
            if in_gamma != out_gamma:
                seq.append(convert(out_order))
            if not out_norm and in_gamma != out_gamma:
                seq.append(ignorm(out_gamma))
            if out_mul:
                seq.append(_scale_tail(out_gamma))
        if not out_norm and out_mul:
            seq.append(_zeroth_mul(out_gamma))

        def apply_seq(x):
            for layer in seq:
                x = layer(x)
            return x

        return Precomputed(layers={"seq": apply_seq})

    @staticmethod
    def _forward(mc: torch.Tensor, *, seq: Callable) -> torch.Tensor:
        if not mc.dtype.is_floating_point:
            mc = mc.to(torch.get_default_dtype())
        return seq(mc)
