"""Short-time Fourier transform as ONE fused kernel (drop-in for diffsptk/modules/stft.py)."""

from __future__ import annotations

import torch

from .. import ops
from ..utils import filter_values, get_layer, pad_mode_id
from .base import BaseFunctionalModule, Precomputed
from .fftr import RealValuedFastFourierTransform
from .frame import Frame
from .spec import Spectrum, spec_format_id
from .window import Window

LEARNABLES = ("basis", "window")


class ShortTimeFourierTransform(BaseFunctionalModule):
    """``(..., T) -> (..., N, L/2+1)``.

    The sub-layers ``frame`` / ``window`` / ``spec`` exist (same attribute and buffer names as the
    reference, e.g. ``stft.window.window``) but ``forward`` does not cascade them: it launches
    ``dsb200_stft`` -- frame, window, real FFT and the spectrum formatter in one pass over HBM.
    """

    def __init__(self, frame_length: int, frame_period: int, fft_length: int, *, center: bool = True,
                 zmean: bool = False, mode: str = "constant", window: str = "blackman", norm: str = "power",
                 symmetric: bool = True, eps: float = 1e-9, relative_floor: float | None = None,
                 out_format: str = "power", learnable: bool | list[str] = False,
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = ShortTimeFourierTransform._precompute(*args, **kwargs, learnable=False, device=x.device,
                                                    dtype=x.dtype, module=False)
        return ShortTimeFourierTransform._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(learnable: bool | list[str]) -> None:
        if isinstance(learnable, (tuple, list)):
            if any(key not in LEARNABLES for key in learnable):
                raise ValueError("An unsupported key is found in learnable.")
        elif not isinstance(learnable, bool):
            raise ValueError("learnable must be boolean or list.")

    @staticmethod
    def _precompute(frame_length: int, frame_period: int, fft_length: int, center: bool, zmean: bool,
                    mode: str, window: str, norm: str, symmetric: bool, eps: float,
                    relative_floor: float | None, out_format: str, learnable: bool | list[str],
                    device: torch.device | None, dtype: torch.dtype | None, module: bool = True) -> Precomputed:
        ShortTimeFourierTransform._check(learnable)
        keys = LEARNABLES if learnable is True else (() if learnable is False else tuple(learnable))
        Frame._check(frame_length, frame_period)
        basis = "basis" in keys and module
        fmt = spec_format_id(out_format, allow_complex=True)
        if fmt == 4:
            RealValuedFastFourierTransform._check(fft_length)
            linear_floor = None
        else:
            Spectrum._check(fft_length, eps, relative_floor)
            linear_floor = None if relative_floor is None else 10 ** (relative_floor / 10)
        values = dict(frame_period=frame_period, fft_length=fft_length, center=center, zmean=zmean,
                      pad_mode=pad_mode_id(mode), eps=eps, relative_floor=linear_floor, out_format=fmt, basis=basis)
        win_params = dict(in_length=frame_length, out_length=fft_length, window=window, norm=norm,
                          symmetric=symmetric, learnable="window" in keys, device=device, dtype=dtype)
        if not module:
            win_params.pop("learnable")
            table = Window._precompute(**win_params).tensors["window"]
            return Precomputed(values=values, tensors={"window_table": table})
        frame = get_layer(True, Frame, dict(frame_length=frame_length, frame_period=frame_period, center=center,
                                            zmean=zmean, mode=mode))
        window_ = get_layer(True, Window, win_params)
        if fmt == 4:
            spec = get_layer(True, RealValuedFastFourierTransform,
                             dict(fft_length=fft_length, out_format="complex", learnable=basis, device=device,
                                  dtype=dtype))
        else:
            spec = get_layer(True, Spectrum, dict(fft_length=fft_length, eps=eps, relative_floor=relative_floor,
                                                  out_format=out_format, learnable=basis))
        return Precomputed(values=values, layers={"frame": frame, "window": window_, "spec": spec})

    @staticmethod
    def _forward(x: torch.Tensor, *, frame_period: int, fft_length: int, center: bool, zmean: bool,
                 pad_mode: int, eps: float, relative_floor: float | None, out_format: int, basis: bool = False,
                 frame=None, window=None, spec=None, window_table: torch.Tensor | None = None) -> torch.Tensor:
        if basis:   # trainable DFT basis: the reference's cascade (stft.py:237-241), every stage a native kernel
            return spec(window(frame(x)))
        table = window_table if window_table is not None else window.window
        y = ops.stft(x, table, frame_period, fft_length, center, zmean, pad_mode, eps,
                     -1.0 if relative_floor is None else relative_floor, out_format)
        return torch.view_as_complex(y) if out_format == 4 else y
