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


_MFCC_TABLES: dict = {}


def _mfcc_tables(x: Tensor, frame_length, fft_length, mfcc_order, n_channel, sample_rate, lifter, window, norm,
                 symmetric, f_min, f_max, scale, erb_factor):
    """Device tables of the default (table-less) call, memoised per parameter set: the functional path of the
    reference rebuilds them on every call; here they -- and the filter-bank plan that hangs on them -- are built
    once per (parameters, device, dtype)."""
    dt = x.dtype if x.dtype.is_floating_point else None
    key = (frame_length, fft_length, mfcc_order, n_channel, sample_rate, lifter, window, norm, symmetric, f_min, f_max,
           scale, erb_factor, x.device, dt)
    t = _MFCC_TABLES.get(key)
    if t is None:
        dev = x.device
        H = tables.make_fbank_matrix(fft_length, n_channel, sample_rate, f_min, f_max, scale, erb_factor, dev, dt)
        cb, ce = tables.column_support(H)
        t = (tables.make_window(frame_length, window, norm, symmetric, device=dev, dtype=dt), H, cb, ce,
             tables.make_dct_matrix(n_channel, 2, dev, dt), tables.make_lifter(mfcc_order, lifter, dev, dt))
        if len(_MFCC_TABLES) > 64:
            _MFCC_TABLES.clear()
        _MFCC_TABLES[key] = t
    return t


def _mfcc_wave_args(x, frame_length, frame_period, fft_length, mfcc_order, n_channel, sample_rate, lifter, center,
                    zmean, mode, window, norm, symmetric, eps, f_min, f_max, floor, gamma, scale, erb_factor,
                    out_format, window_table, H, W, liftering_vector, H_begin, H_end):
    if window_table is None and H is None and W is None and liftering_vector is None:
        window_table, H, cb, ce, W, liftering_vector = _mfcc_tables(
            x, frame_length, fft_length, mfcc_order, n_channel, sample_rate, lifter, window, norm, symmetric,
            f_min, f_max, scale, erb_factor)
    else:
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
        cb, ce = support_of(H, H_begin, H_end)
    if H.shape[0] != fft_length // 2 + 1:   # mfcc.py:243 -> fbank.py:305: the spectrum must match the filter bank
        raise ValueError(f"dimension of spectrum must be {H.shape[0]}, but got {fft_length // 2 + 1}.")
    if W.shape[0] != H.shape[1] or liftering_vector.shape[-1] > H.shape[1]:
        raise ValueError("the DCT matrix / liftering vector do not match the filter bank.")
    plan = ops.mfcc_plan(cb, ce, H.shape[0]) if x.dtype != torch.float64 else None
    return (window_table, H, cb, ce, W, liftering_vector, frame_period, fft_length, center, zmean, pad_mode_id(mode),
            eps, floor, gamma, mfcc_format_id(out_format), plan)


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
    a = _mfcc_wave_args(x, frame_length, frame_period, fft_length, mfcc_order, n_channel, sample_rate, lifter, center,
                        zmean, mode, window, norm, symmetric, eps, f_min, f_max, floor, gamma, scale, erb_factor,
                        out_format, window_table, H, W, liftering_vector, H_begin, H_end)
    try:
        return ops.mfcc_wave(x, *a)
    except NotImplementedError:
        # configuration outside the single-kernel envelope: two kernels, spectrum round-trips through HBM
        (window_table, H, cb, ce, W, liftering_vector, frame_period, fft_length, center, zmean, pad, eps, floor, gamma,
         fmt, _) = a
        P = ops.stft(x, window_table, frame_period, fft_length, center, zmean, pad, eps, -1.0, 3)
        return ops.mfcc(P, H, cb, ce, W, liftering_vector, floor, gamma, fmt)


def mfcc_from_waveform_gather(x: Tensor, out: Tensor, dst_ptrs, mc_ptr: int, rank: int, *, frame_length: int = 400,
                              frame_period: int = 80, fft_length: int = 512, mfcc_order: int = 13,
                              n_channel: int = 40, sample_rate: int = 16000, out_format: str | int = "y") -> Tensor:
    """``mfcc_from_waveform`` of this rank's ``[B_local, T]`` waveforms with the all-gather fused into the kernel:
    the rows are stored at rank ``rank``'s slab of ``out`` (``[world * B_local, N, D]``, symmetric memory) on every
    rank -- through ``mc_ptr`` (the NVSwitch multicast address of ``out``) when non-zero, else through
    ``dst_ptrs`` (the address of ``out`` in every rank's memory).  See ``distributed.FusedGatherMfcc``."""
    a = _mfcc_wave_args(x, frame_length, frame_period, fft_length, mfcc_order, n_channel, sample_rate, 1, True,
                        False, "constant", "blackman", "power", True, 1e-9, 0, None, 1e-5, 0, "htk", None,
                        out_format, None, None, None, None, None, None)
    n = ops.num_frames(x.shape[-1], frame_period)
    dst = [mc_ptr] if mc_ptr else list(dst_ptrs)
    return ops.mfcc_wave_gather(x, out, dst, rank * x.shape[0] * n, *a)


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
