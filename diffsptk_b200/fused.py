"""Fused waveform-to-feature pipelines of BASELINE.json.

``lpc_from_waveform`` = ``LPC(Window(Frame(x)))`` (README.md:198-201 of the reference) and
``mfcc_from_waveform`` = ``MFCC(STFT(x))`` without materialising the framed / windowed / spectral
intermediates in HBM.  ``fuse()`` recognises those cascades in an ``nn.Sequential`` built from this
package's modules and returns the fused equivalent.
"""

from __future__ import annotations

import torch
from torch import Tensor, nn

from . import ops, tables
from .modules.fbank import support_of
from .modules.levdur import default_eps
from .modules.mfcc import mfcc_format_id
from .utils import pad_mode_id


def lpc_from_waveform(x: Tensor, *, frame_length: int = 400, frame_period: int = 80, lpc_order: int = 24,
                      center: bool = True, zmean: bool = False, mode: str = "constant",
                      window: str = "blackman", norm: str = "power", symmetric: bool = True,
                      eps: float | None = None, window_table: Tensor | None = None) -> Tensor:
    """``(..., T) -> (..., N, M+1)``: frame -> window -> autocorrelation -> Levinson-Durbin."""
    if frame_length <= lpc_order:
        raise ValueError("acr_order must be less than frame_length.")
    dt = x.dtype if x.dtype.is_floating_point else None
    if window_table is None:
        window_table = tables.make_window(frame_length, window, norm, symmetric, device=x.device, dtype=dt)
    return ops.lpc_wave(x, window_table, frame_period, center, zmean, pad_mode_id(mode), lpc_order,
                        float(default_eps(eps, dt)))


def mfcc_from_waveform(x: Tensor, *, frame_length: int = 400, frame_period: int = 80, fft_length: int = 512,
                       mfcc_order: int = 13, n_channel: int = 40, sample_rate: int = 16000, lifter: int = 1,
                       center: bool = True, zmean: bool = False, mode: str = "constant",
                       window: str = "blackman", norm: str = "power", symmetric: bool = True,
                       eps: float = 1e-9, f_min: float = 0, f_max: float | None = None, floor: float = 1e-5,
                       gamma: float = 0, scale: str = "htk", erb_factor: float | None = None,
                       out_format: str | int = "y", window_table: Tensor | None = None,
                       H: Tensor | None = None, W: Tensor | None = None,
                       liftering_vector: Tensor | None = None, H_begin: Tensor | None = None,
                       H_end: Tensor | None = None) -> Tensor:
    """``(..., T) -> (..., N, D)``: STFT power (``eps``, no relative floor) -> fbank -> DCT -> lifter."""
    dt = x.dtype if x.dtype.is_floating_point else None
    dev = x.device
    if window_table is None:
        window_table = tables.make_window(frame_length, window, norm, symmetric, device=dev, dtype=dt)
    if H is None:
        H = tables.make_fbank_matrix(fft_length, n_channel, sample_rate, f_min, f_max, scale, erb_factor, dev, dt)
    if W is None:
        W = tables.make_dct_matrix(H.shape[1], 2, dev, dt)
    if liftering_vector is None:
        liftering_vector = tables.make_lifter(mfcc_order, lifter, dev, dt)
    fmt = mfcc_format_id(out_format)
    cb, ce = support_of(H, H_begin, H_end)
    try:
        return ops.mfcc_wave(x, window_table, H, cb, ce, W, liftering_vector, frame_period, fft_length, center,
                             zmean, pad_mode_id(mode), eps, floor, gamma, fmt)
    except NotImplementedError:
        # configuration outside the single-kernel envelope: two kernels, spectrum round-trips through HBM
        P = ops.stft(x, window_table, frame_period, fft_length, center, zmean, pad_mode_id(mode), eps, -1.0, 3)
        return ops.mfcc(P, H, cb, ce, W, liftering_vector, floor, gamma, fmt)


class FusedLPC(nn.Module):
    """Fused equivalent of ``Sequential(Frame, Window, LPC)``."""

    def __init__(self, frame, window, lpc):
        super().__init__()
        self.frame, self.window, self.lpc = frame, window, lpc

    def forward(self, x: Tensor) -> Tensor:
        f, l = self.frame, self.lpc
        return lpc_from_waveform(x, frame_length=f.frame_length, frame_period=f.frame_period,
                                 lpc_order=l.lpc_order, center=f.center, zmean=f.zmean, mode=f.mode, eps=l.eps,
                                 window_table=self.window.window)


class FusedMFCC(nn.Module):
    """Fused equivalent of ``Sequential(STFT(out_format='power'), MFCC)``."""

    def __init__(self, stft, mfcc):
        super().__init__()
        self.stft, self.mfcc = stft, mfcc

    def forward(self, x: Tensor) -> Tensor:
        s, m = self.stft, self.mfcc
        modes = {0: "constant", 1: "reflect", 2: "replicate", 3: "circular"}
        return mfcc_from_waveform(
            x, frame_period=s.frame_period, fft_length=s.fft_length, center=s.center, zmean=s.zmean,
            mode=modes[s.pad_mode], eps=s.eps, floor=m.floor, gamma=m.gamma, out_format=m.out_format,
            window_table=s.window.window, H=m.fbank.H, W=m.dct.W, liftering_vector=m.liftering_vector,
            H_begin=getattr(m.fbank, "H_begin", None), H_end=getattr(m.fbank, "H_end", None))


def fuse(seq: nn.Sequential) -> nn.Module:
    """Replace a recognised cascade by its fused kernel chain; anything else is returned unchanged."""
    from . import modules as M

    layers = list(seq)
    if (len(layers) == 3 and isinstance(layers[0], M.Frame) and isinstance(layers[1], M.Window)
            and isinstance(layers[2], M.LPC) and layers[1].out_length in (None, layers[1].in_dim)):
        return FusedLPC(*layers)
    if (len(layers) == 2 and isinstance(layers[0], M.STFT) and isinstance(layers[1], M.MFCC)
            and layers[0].out_format == 3 and layers[0].relative_floor is None):
        return FusedMFCC(*layers)
    return seq
