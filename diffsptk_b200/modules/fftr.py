"""Real FFT with output formatter (drop-in for diffsptk/modules/fftr.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import filter_values
from .base import BaseFunctionalModule, Precomputed

_FORMATS = {"complex": 0, "real": 1, "imaginary": 2, "amplitude": 3, "power": 4}


def _format_id(out_format) -> int:
    if isinstance(out_format, str) and out_format in _FORMATS:
        return _FORMATS[out_format]
    if isinstance(out_format, int) and not isinstance(out_format, bool) and 0 <= out_format <= 4:
        return out_format
    raise ValueError(f"out_format {out_format} is not supported.")


class RealValuedFastFourierTransform(BaseFunctionalModule):
    """``(..., N) -> (..., L/2+1)``; kernel ``dsb200_rfft`` (no cuFFT / torch.fft on the path)."""

    def __init__(self, fft_length: int, out_format: str | int = "complex", learnable: bool = False,
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())), learnable=learnable is True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = RealValuedFastFourierTransform._precompute(*args, **kwargs, learnable=False, device=x.device,
                                                         dtype=x.dtype)
        return RealValuedFastFourierTransform._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(fft_length: int | None) -> None:
        if fft_length is not None and (fft_length <= 0 or fft_length % 2 == 1):
            raise ValueError("fft_length must be positive even.")

    @staticmethod
    def _precompute(fft_length: int | None, out_format: str | int, learnable: bool,
                    device: torch.device | None, dtype: torch.dtype | None) -> Precomputed:
        RealValuedFastFourierTransform._check(fft_length)
        fmt = _format_id(out_format)
        tensors = {}
        if learnable:
            # The reference switches to a DFT-by-matmul with a trainable basis (fftr.py:123-131,146-150): W is
            # [fft_length, 2 (L/2+1)] = (cos | -sin), built in double like the reference's fft(eye) and cast.
            if fft_length is None:
                raise ValueError("fft_length must be specified when learnable is True.")
            W = torch.fft.fft(torch.eye(fft_length, dtype=torch.double))[..., : fft_length // 2 + 1]
            W = torch.cat([W.real, W.imag], dim=-1)
            tensors["W"] = W.to(device=device, dtype=dtype if dtype is not None and dtype.is_floating_point
                                else torch.get_default_dtype())
        return Precomputed(values={"fft_length": fft_length, "out_format": fmt}, tensors=tensors)

    @staticmethod
    def _forward(x: torch.Tensor, *, fft_length: int | None, out_format: int,
                 W: torch.Tensor | None = None) -> torch.Tensor:
        n = x.size(-1) if fft_length is None else fft_length
        if n % 2 == 1:
            raise ValueError("fft_length must be positive even.")
        if W is not None:
            # trainable basis: a dense product on the native row-times-matrix kernel (differentiable in x and W),
            # then the formatter as element-wise torch ops
            if x.size(-1) != n:
                x = torch.nn.functional.pad(x, (0, n - x.size(-1)))
            re, im = torch.tensor_split(ops.rowmat(x, W), 2, dim=-1)
            if out_format == 0:
                return torch.complex(re, im)
            if out_format == 1:
                return re
            if out_format == 2:
                return im
            p2 = re * re + im * im
            return torch.sqrt(p2) if out_format == 3 else p2
        y = ops.rfft(x, n, out_format)
        return torch.view_as_complex(y) if out_format == 0 else y
