"""Spectrum of b / a (drop-in for diffsptk/modules/spec.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import filter_values
from .base import BaseFunctionalModule, Precomputed

_FORMATS = {"db": 0, "log-magnitude": 1, "magnitude": 2, "power": 3}


def spec_format_id(out_format, allow_complex: bool = False) -> int:
    if isinstance(out_format, str):
        if out_format in _FORMATS:
            return _FORMATS[out_format]
        if allow_complex and out_format == "complex":
            return 4
    elif isinstance(out_format, int) and not isinstance(out_format, bool) and 0 <= out_format <= 3:
        return out_format
    raise ValueError(f"out_format {out_format} is not supported.")


class Spectrum(BaseFunctionalModule):
    """``(..., M+1) [, (..., N+1)] -> (..., L/2+1)``; kernel ``dsb200_spec``."""

    def __init__(self, fft_length: int, *, eps: float = 0, relative_floor: float | None = None,
                 out_format: str | int = "power", learnable: bool = False) -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, b: torch.Tensor | None = None, a: torch.Tensor | None = None) -> torch.Tensor:
        return self._call_forward(b, a)

    @staticmethod
    def _func(b: torch.Tensor | None = None, a: torch.Tensor | None = None, *args, **kwargs) -> torch.Tensor:
        pre = Spectrum._precompute(*args, **kwargs, module=False)
        return Spectrum._apply_precomputed(pre, b=b, a=a)

    @staticmethod
    def _check(fft_length: int, eps: float, relative_floor: float | None) -> None:
        if fft_length <= 1:
            raise ValueError("fft_length must be greater than 1.")
        if fft_length % 2 == 1:
            raise ValueError("fft_length must be positive even.")
        if eps < 0:
            raise ValueError("eps must be non-negative.")
        if relative_floor is not None and 0 <= relative_floor:
            raise ValueError("relative_floor must be negative.")

    @staticmethod
    def _precompute(fft_length: int, eps: float, relative_floor: float | None, out_format: str | int,
                    learnable: bool = False, module: bool = True) -> Precomputed:
        Spectrum._check(fft_length, eps, relative_floor)
        linear_floor = None if relative_floor is None else 10 ** (relative_floor / 10)  # spec.py:121-122
        layers = {}
        if learnable and module:
            # trainable DFT basis (spec.py:134-142 -> fftr.py:123-131): the amplitude comes from the sub-layer's dense
            # product instead of the fused kernel; same sub-layer name as the reference (``spec.fftr.W``)
            from .fftr import RealValuedFastFourierTransform
            layers["fftr"] = RealValuedFastFourierTransform(fft_length, out_format="amplitude", learnable=True)
        return Precomputed(values={"fft_length": fft_length, "eps": eps, "relative_floor": linear_floor,
                                   "out_format": spec_format_id(out_format)}, layers=layers)

    @staticmethod
    def _forward(b: torch.Tensor | None, a: torch.Tensor | None, *, fft_length: int, eps: float,
                 relative_floor: float | None, out_format: int, fftr=None) -> torch.Tensor:
        if b is None and a is None:
            raise ValueError("Either b or a must be specified.")
        if fftr is not None:   # trainable basis: the reference's composite, spec.py:162-178
            if a is not None:
                K, a1 = torch.split(a, [1, a.size(-1) - 1], dim=-1)
                a1 = torch.nn.functional.pad(a1, (1, 0), value=1.0)
                X = K * (fftr(b) / fftr(a1)) if b is not None else K / fftr(a1)
            else:
                X = fftr(b)
            return _format_power(torch.square(X) + eps, relative_floor, out_format)
        rf = -1.0 if relative_floor is None else relative_floor
        needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (b, a))
        if a is not None and needs_grad:
            return _polezero_differentiable(b, a, fft_length, eps, relative_floor, out_format)
        return ops.spec(b, a, fft_length, eps, rf, out_format)


def _polezero_differentiable(b, a, fft_length, eps, relative_floor, out_format):
    """``K |B| / |A|`` spectra under autograd (spec.py:162-178 of the reference): the fused kernel of the forward-only
    path differentiates the numerator only, so with a denominator that (or whose partner) needs a gradient the two
    power spectra come from the differentiable numerator kernel (``|B|^2``, ``|A|^2`` with the gain removed) and the
    ratio, floor and formatter are element-wise torch ops -- gradients reach both ``b`` and ``a``."""
    K, a1 = torch.split(a, [1, a.size(-1) - 1], dim=-1)
    a1 = torch.nn.functional.pad(a1, (1, 0), value=1.0)                      # utils/private.py:200-209 remove_gain
    s = torch.square(K) / ops.spec(a1, None, fft_length, 0.0, -1.0, 3)
    if b is not None:
        s = s * ops.spec(b, None, fft_length, 0.0, -1.0, 3)
    return _format_power(s + eps, relative_floor, out_format)


def _format_power(s, relative_floor, out_format):
    """Relative floor and formatter of spec.py:173-177 as element-wise torch ops (the differentiable composites)."""
    if relative_floor is not None:
        s = torch.maximum(s, torch.amax(s, dim=-1, keepdim=True) * relative_floor)
    if out_format == 0:
        return 10 * torch.log10(s)
    if out_format == 1:
        return 0.5 * torch.log(s)
    if out_format == 2:
        return torch.sqrt(s)
    return s
