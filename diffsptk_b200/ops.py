"""``torch.ops.diffsptk_b200.*``: the boundary between the nn.Module mirror and the C ABI.

Each op flattens leading batch dims, allocates the output with torch (the
library owns no data memory), and calls the matching ``dsb200_*`` entry point
with raw device pointers and the current CUDA stream.  Ops are registered for
CUDA only: a CPU tensor fails in the dispatcher -- there is no fallback path.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
from torch import Tensor

from . import _native as N

_NS = "diffsptk_b200"


# ------------------------------------------------------------------------------------ helpers
def _f64(t: Tensor) -> bool:
    return t.dtype == torch.float64


def _native_dtype(*ts: Tensor) -> torch.dtype:
    """float64 if any operand is float64, else float32 (ints / half / bf16 are up-cast)."""
    return torch.float64 if any(t is not None and t.dtype == torch.float64 for t in ts) else torch.float32


def _prep(t: Optional[Tensor], dtype: torch.dtype) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _ptr(t: Optional[Tensor]):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


def _ptr_or_dummy(t: Tensor):
    # zero-sized tensors have no storage; the library never dereferences when rows == 0
    return C.c_void_p(t.data_ptr()) if t.numel() else None


def _stream(t: Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _dev(t: Tensor) -> int:
    if not t.is_cuda:
        raise RuntimeError("diffsptk_b200 ops run on CUDA tensors only (no CPU fallback).")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _no_grad_check(*ts):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts):
        raise NotImplementedError(
            "this diffsptk_b200 op is forward-only for that input (differentiable: every op with respect to its signal "
            "input, learnable windows, filter banks and DFT bases included): wrap the call in torch.no_grad() or "
            "detach() the inputs."
        )


def num_frames(T: int, frame_period: int) -> int:
    return 0 if T <= 0 else (T - 1) // frame_period + 1


def _frame_params(L, P, center, zmean, pad_mode) -> N.FrameParams:
    return N.FrameParams(int(L), int(P), int(bool(center)), int(bool(zmean)), int(pad_mode))


def _spec_params(fft_length, out_format, eps, relative_floor) -> N.SpecParams:
    has = relative_floor is not None and relative_floor > 0
    return N.SpecParams(int(fft_length), int(out_format), int(has), 0, float(eps), float(relative_floor) if has else 0.0)


# ------------------------------------------------------------------------------------- kernels
@torch.library.custom_op(f"{_NS}::frame", mutates_args=(), device_types="cuda")
def frame(x: Tensor, frame_length: int, frame_period: int, center: bool, zmean: bool, pad_mode: int) -> Tensor:
    dt = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32
    xc = _prep(x, dt)
    T = xc.shape[-1]
    if T < 1:
        raise ValueError("waveform length must be at least 1")
    lead = xc.shape[:-1]
    B = xc.numel() // T
    n = num_frames(T, frame_period)
    y = torch.empty((*lead, n, frame_length), device=x.device, dtype=dt)
    p = _frame_params(frame_length, frame_period, center, zmean, pad_mode)
    N.check(N.typed("dsb200_frame", dt == torch.float64)(_ptr(xc), _ptr(y), B, T, C.byref(p), _dev(x), _stream(x)))
    return y


@frame.register_fake
def _(x, frame_length, frame_period, center, zmean, pad_mode):
    dt = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32
    return x.new_empty((*x.shape[:-1], num_frames(x.shape[-1], frame_period), frame_length), dtype=dt)


@torch.library.custom_op(f"{_NS}::window", mutates_args=(), device_types="cuda")
def window(x: Tensor, w: Tensor, out_length: int) -> Tensor:
    dt = _native_dtype(x, w)
    xc, wc = _prep(x, dt), _prep(w, dt)
    L1 = xc.shape[-1]
    rows = xc.numel() // max(L1, 1)
    y = torch.empty((*xc.shape[:-1], out_length), device=x.device, dtype=dt)
    N.check(N.typed("dsb200_window", dt == torch.float64)(_ptr(xc), _ptr(wc), _ptr(y), rows, L1, out_length, _dev(x), _stream(x)))
    return y


@window.register_fake
def _(x, w, out_length):
    return x.new_empty((*x.shape[:-1], out_length), dtype=_native_dtype(x, w))


@torch.library.custom_op(f"{_NS}::rfft", mutates_args=(), device_types="cuda")
def rfft(x: Tensor, fft_length: int, out_format: int) -> Tensor:
    """Returns real [..., K] or, for out_format 0, real [..., K, 2] (view_as_complex by the caller)."""
    dt = _native_dtype(x)
    xc = _prep(x, dt)
    Lin = xc.shape[-1]
    rows = xc.numel() // max(Lin, 1)
    K = fft_length // 2 + 1
    shape = (*xc.shape[:-1], K, 2) if out_format == 0 else (*xc.shape[:-1], K)
    y = torch.empty(shape, device=x.device, dtype=dt)
    N.check(N.typed("dsb200_rfft", dt == torch.float64)(_ptr(xc), _ptr(y), rows, Lin, fft_length, out_format, _dev(x), _stream(x)))
    return y


@rfft.register_fake
def _(x, fft_length, out_format):
    K = fft_length // 2 + 1
    shape = (*x.shape[:-1], K, 2) if out_format == 0 else (*x.shape[:-1], K)
    return x.new_empty(shape, dtype=_native_dtype(x))


@torch.library.custom_op(f"{_NS}::spec", mutates_args=(), device_types="cuda")
def spec(b: Optional[Tensor], a: Optional[Tensor], fft_length: int, eps: float, relative_floor: float,
         out_format: int) -> Tensor:
    """relative_floor is the LINEAR factor (<= 0 means none)."""
    ref = b if b is not None else a
    if ref is None:
        raise ValueError("Either b or a must be specified.")
    dt = _native_dtype(b, a)
    bc, ac = _prep(b, dt), _prep(a, dt)
    lead = ref.shape[:-1]
    if bc is not None and ac is not None and bc.shape[:-1] != ac.shape[:-1]:
        lead = torch.broadcast_shapes(bc.shape[:-1], ac.shape[:-1])
        bc = bc.expand(*lead, bc.shape[-1]).contiguous()
        ac = ac.expand(*lead, ac.shape[-1]).contiguous()
    rows = 1
    for s in lead:
        rows *= s
    K = fft_length // 2 + 1
    y = torch.empty((*lead, K), device=ref.device, dtype=dt)
    p = _spec_params(fft_length, out_format, eps, relative_floor)
    N.check(N.typed("dsb200_spec", dt == torch.float64)(
        _ptr(bc), 0 if bc is None else bc.shape[-1], _ptr(ac), 0 if ac is None else ac.shape[-1],
        _ptr(y), rows, C.byref(p), _dev(ref), _stream(ref)))
    return y


@spec.register_fake
def _(b, a, fft_length, eps, relative_floor, out_format):
    ref = b if b is not None else a
    lead = ref.shape[:-1]
    if b is not None and a is not None:
        lead = torch.broadcast_shapes(b.shape[:-1], a.shape[:-1])
    return ref.new_empty((*lead, fft_length // 2 + 1), dtype=_native_dtype(b, a))


@torch.library.custom_op(f"{_NS}::stft", mutates_args=(), device_types="cuda")
def stft(x: Tensor, window: Tensor, frame_period: int, fft_length: int, center: bool, zmean: bool,
         pad_mode: int, eps: float, relative_floor: float, out_format: int) -> Tensor:
    """Fused frame+window+rFFT+formatter.  Complex output comes back as real [..., N, K, 2]."""
    dt = _native_dtype(x, window)
    xc, wc = _prep(x, dt), _prep(window, dt)
    T = xc.shape[-1]
    if T < 1:
        raise ValueError("waveform length must be at least 1")
    B = xc.numel() // T
    L = wc.shape[-1]
    n = num_frames(T, frame_period)
    K = fft_length // 2 + 1
    shape = (*xc.shape[:-1], n, K, 2) if out_format == 4 else (*xc.shape[:-1], n, K)
    y = torch.empty(shape, device=x.device, dtype=dt)
    p = N.StftParams(_frame_params(L, frame_period, center, zmean, pad_mode),
                     _spec_params(fft_length, out_format, eps, relative_floor))
    N.check(N.typed("dsb200_stft", dt == torch.float64)(_ptr(xc), _ptr(wc), _ptr(y), B, T, C.byref(p), _dev(x), _stream(x)))
    return y


@stft.register_fake
def _(x, window, frame_period, fft_length, center, zmean, pad_mode, eps, relative_floor, out_format):
    n, K = num_frames(x.shape[-1], frame_period), fft_length // 2 + 1
    shape = (*x.shape[:-1], n, K, 2) if out_format == 4 else (*x.shape[:-1], n, K)
    return x.new_empty(shape, dtype=_native_dtype(x, window))


@torch.library.custom_op(f"{_NS}::acorr", mutates_args=(), device_types="cuda")
def acorr(x: Tensor, acr_order: int, out_format: int) -> Tensor:
    dt = _native_dtype(x)
    xc = _prep(x, dt)
    L = xc.shape[-1]
    rows = xc.numel() // max(L, 1)
    r = torch.empty((*xc.shape[:-1], acr_order + 1), device=x.device, dtype=dt)
    N.check(N.typed("dsb200_acorr", dt == torch.float64)(_ptr(xc), _ptr(r), rows, L, acr_order, out_format, _dev(x), _stream(x)))
    return r


@acorr.register_fake
def _(x, acr_order, out_format):
    return x.new_empty((*x.shape[:-1], acr_order + 1), dtype=_native_dtype(x))


@torch.library.custom_op(f"{_NS}::levdur", mutates_args=(), device_types="cuda")
def levdur(r: Tensor, eps: float) -> Tensor:
    dt = _native_dtype(r)
    rc = _prep(r, dt)
    D = rc.shape[-1]
    rows = rc.numel() // max(D, 1)
    a = torch.empty_like(rc)
    N.check(N.typed("dsb200_levdur", dt == torch.float64)(_ptr(rc), _ptr(a), rows, D - 1, eps, _dev(r), _stream(r)))
    return a


@levdur.register_fake
def _(r, eps):
    return r.new_empty(r.shape, dtype=_native_dtype(r))


@torch.library.custom_op(f"{_NS}::lpc", mutates_args=(), device_types="cuda")
def lpc(x: Tensor, lpc_order: int, eps: float) -> Tensor:
    dt = _native_dtype(x)
    xc = _prep(x, dt)
    L = xc.shape[-1]
    rows = xc.numel() // max(L, 1)
    a = torch.empty((*xc.shape[:-1], lpc_order + 1), device=x.device, dtype=dt)
    N.check(N.typed("dsb200_lpc", dt == torch.float64)(_ptr(xc), _ptr(a), rows, L, lpc_order, eps, _dev(x), _stream(x)))
    return a


@lpc.register_fake
def _(x, lpc_order, eps):
    return x.new_empty((*x.shape[:-1], lpc_order + 1), dtype=_native_dtype(x))


@torch.library.custom_op(f"{_NS}::lpc_wave", mutates_args=(), device_types="cuda")
def lpc_wave(x: Tensor, window: Tensor, frame_period: int, center: bool, zmean: bool, pad_mode: int,
             lpc_order: int, eps: float) -> Tensor:
    dt = _native_dtype(x, window)
    xc, wc = _prep(x, dt), _prep(window, dt)
    T = xc.shape[-1]
    B = xc.numel() // max(T, 1)
    n = num_frames(T, frame_period)
    a = torch.empty((*xc.shape[:-1], n, lpc_order + 1), device=x.device, dtype=dt)
    p = _frame_params(wc.shape[-1], frame_period, center, zmean, pad_mode)
    N.check(N.typed("dsb200_lpc_wave", dt == torch.float64)(_ptr(xc), _ptr(wc), _ptr(a), B, T, C.byref(p), lpc_order, eps, _dev(x), _stream(x)))
    return a


@lpc_wave.register_fake
def _(x, window, frame_period, center, zmean, pad_mode, lpc_order, eps):
    return x.new_empty((*x.shape[:-1], num_frames(x.shape[-1], frame_period), lpc_order + 1), dtype=_native_dtype(x, window))


@torch.library.custom_op(f"{_NS}::rowmat", mutates_args=(), device_types="cuda")
def rowmat(x: Tensor, W: Tensor) -> Tensor:
    dt = _native_dtype(x, W)
    xc, Wc = _prep(x, dt), _prep(W, dt)
    Din, Dout = Wc.shape
    rows = xc.numel() // max(Din, 1)
    y = torch.empty((*xc.shape[:-1], Dout), device=x.device, dtype=dt)
    N.check(N.typed("dsb200_rowmat", dt == torch.float64)(_ptr(xc), _ptr(Wc), _ptr(y), rows, Din, Dout, _dev(x), _stream(x)))
    return y


@rowmat.register_fake
def _(x, W):
    return x.new_empty((*x.shape[:-1], W.shape[1]), dtype=_native_dtype(x, W))


@torch.library.custom_op(f"{_NS}::mcep", mutates_args=(), device_types="cuda")
def mcep(x: Tensor, P0: Tensor, G: Tensor, Hm: Tensor, alpha_vector: Tensor, n_iter: int) -> Tensor:
    dt = _native_dtype(x, alpha_vector)
    xc = _prep(x, dt)
    K = xc.shape[-1]
    D = alpha_vector.shape[-1]
    rows = xc.numel() // max(K, 1)
    mc = torch.empty((*xc.shape[:-1], D), device=x.device, dtype=dt)
    p = N.McepParams(2 * (K - 1), D - 1, n_iter, 0)
    N.check(N.typed("dsb200_mcep", dt == torch.float64)(
        _ptr(xc), _ptr(mc), rows, C.byref(p), _ptr(_prep(P0, dt)), _ptr(_prep(G, dt)), _ptr(_prep(Hm, dt)),
        _ptr(_prep(alpha_vector, dt)), _dev(x), _stream(x)))
    return mc


@mcep.register_fake
def _(x, P0, G, Hm, alpha_vector, n_iter):
    return x.new_empty((*x.shape[:-1], alpha_vector.shape[-1]), dtype=_native_dtype(x, alpha_vector))


@torch.library.custom_op(f"{_NS}::fbank", mutates_args=(), device_types="cuda")
def fbank(x: Tensor, H: Tensor, col_begin: Optional[Tensor], col_end: Optional[Tensor], floor: float,
          gamma: float, use_power: bool, want_energy: bool) -> tuple[Tensor, Tensor]:
    dt = _native_dtype(x, H)
    xc, Hc = _prep(x, dt), _prep(H, dt)
    K, Cn = Hc.shape
    rows = xc.numel() // max(K, 1)
    y = torch.empty((*xc.shape[:-1], Cn), device=x.device, dtype=dt)
    E = torch.empty((*xc.shape[:-1], 1) if want_energy else (0,), device=x.device, dtype=dt)
    p = N.FbankParams(2 * (K - 1), Cn, int(use_power), int(want_energy), float(floor), float(gamma))
    N.check(N.typed("dsb200_fbank", dt == torch.float64)(
        _ptr(xc), _ptr(Hc), _ptr(col_begin), _ptr(col_end), _ptr(y), _ptr(E) if want_energy else None, rows,
        C.byref(p), _dev(x), _stream(x)))
    return y, E


@fbank.register_fake
def _(x, H, col_begin, col_end, floor, gamma, use_power, want_energy):
    dt = _native_dtype(x, H)
    return (x.new_empty((*x.shape[:-1], H.shape[1]), dtype=dt),
            x.new_empty((*x.shape[:-1], 1) if want_energy else (0,), dtype=dt))


def _mfcc_dim(M: int, out_format: int) -> int:
    return M + (0, 1, 1, 2)[out_format]


def _check_mfcc_tables(K: int, Cn: int, W: Tensor, M: int) -> None:
    """Table shapes the kernels index without bounds checks: ``H [K, C]``, ``W [C, C]``, lifter ``[M + 1]``."""
    if W.dim() != 2 or W.shape[0] != Cn or W.shape[1] != Cn:   # the kernels read W with row pitch C (dct.py:135-137)
        raise ValueError(f"DCT matrix must be [{Cn}, {Cn}] for a {Cn}-channel filter bank, got {tuple(W.shape)}.")
    if M + 1 > Cn:   # mfcc.py:165-166
        raise ValueError("mfcc_order must be less than n_channel.")


@torch.library.custom_op(f"{_NS}::mfcc", mutates_args=(), device_types="cuda")
def mfcc(x: Tensor, H: Tensor, col_begin: Optional[Tensor], col_end: Optional[Tensor], W: Tensor,
         lifter: Tensor, floor: float, gamma: float, out_format: int) -> Tensor:
    dt = _native_dtype(x, H)
    xc, Hc, Wc, lc = _prep(x, dt), _prep(H, dt), _prep(W, dt), _prep(lifter, dt)
    K, Cn = Hc.shape
    M = lc.shape[-1] - 1
    _check_mfcc_tables(K, Cn, Wc, M)
    if xc.shape[-1] != K:   # mfcc.py:243 -> fbank.py:305 check_size
        raise ValueError(f"dimension of input must be {K}, but got {xc.shape[-1]}.")
    rows = xc.numel() // max(K, 1)
    y = torch.empty((*xc.shape[:-1], _mfcc_dim(M, out_format)), device=x.device, dtype=dt)
    p = N.MfccParams(N.FbankParams(2 * (K - 1), Cn, 0, 0, float(floor), float(gamma)), M, out_format)
    N.check(N.typed("dsb200_mfcc", dt == torch.float64)(
        _ptr(xc), _ptr(Hc), _ptr(col_begin), _ptr(col_end), _ptr(Wc), _ptr(lc), _ptr(y), rows, C.byref(p),
        _dev(x), _stream(x)))
    return y


@mfcc.register_fake
def _(x, H, col_begin, col_end, W, lifter, floor, gamma, out_format):
    return x.new_empty((*x.shape[:-1], _mfcc_dim(lifter.shape[-1] - 1, out_format)), dtype=_native_dtype(x, H))


MAX_GATHER_PEERS = 8   # destinations of one feature row in dsb200_mfcc_wave_ex


def mfcc_plan(col_begin: Optional[Tensor], col_end: Optional[Tensor], n_bins: int) -> Optional[Tensor]:
    """Segment plan of the fused MFCC kernel's filter-bank stage (``dsb200_mfcc_plan_build``): built on the host
    from the filter supports (one device->host copy of 2 C integers), cached on the ``col_begin`` tensor object, so
    a module or a memoised functional table pays for it once.  None when there is no support or no plan."""
    if col_begin is None or col_end is None or os.environ.get("DSB200_MFCC_PLAN") == "0":   # knob: A/B timing
        return None
    cached = getattr(col_begin, "_dsb200_plan", None)
    if cached is not None and cached[0] is col_end:
        return cached[1]
    cb = col_begin.detach().to("cpu", torch.int32).contiguous()
    ce = col_end.detach().to("cpu", torch.int32).contiguous()
    Cn = cb.numel()
    lib = N.load()
    plan = torch.zeros(int(lib.dsb200_mfcc_plan_ints(Cn)), dtype=torch.int32)
    i32p = C.POINTER(C.c_int32)
    N.check(lib.dsb200_mfcc_plan_build(C.cast(cb.data_ptr(), i32p), C.cast(ce.data_ptr(), i32p), Cn, int(n_bins),
                                       C.cast(plan.data_ptr(), i32p)))
    out = plan.to(col_begin.device) if int(plan[0]) > 0 else None
    try:
        col_begin._dsb200_plan = (col_end, out)
    except Exception:
        pass
    return out


def _mfcc_wave_call(x, window, H, col_begin, col_end, W, lifter, plan, dst_ptrs, row_offset, frame_period,
                    fft_length, center, zmean, pad_mode, eps, floor, gamma, out_format):
    """Shared launcher of ``mfcc_wave`` (one destination) and ``mfcc_wave_gather`` (one per rank)."""
    dt = _native_dtype(x, window, H)
    xc, wc, Hc, Wc, lc = (_prep(t, dt) for t in (x, window, H, W, lifter))
    T = xc.shape[-1]
    B = xc.numel() // max(T, 1)
    K, Cn = Hc.shape
    M = lc.shape[-1] - 1
    _check_mfcc_tables(K, Cn, Wc, M)
    if K != fft_length // 2 + 1:   # the reference's MFCC raises when its fft_length differs from the STFT's (fbank.py:305)
        raise ValueError(f"dimension of input must be {K}, but got {fft_length // 2 + 1}.")
    sp = N.StftParams(_frame_params(wc.shape[-1], frame_period, center, zmean, pad_mode),
                      _spec_params(fft_length, 3, eps, None))
    mp = N.MfccParams(N.FbankParams(fft_length, Cn, 0, 0, float(floor), float(gamma)), M, out_format)
    dst = (C.c_void_p * len(dst_ptrs))(*dst_ptrs)
    N.check(N.typed("dsb200_mfcc_wave_ex", dt == torch.float64)(
        _ptr(xc), _ptr(wc), _ptr(Hc), _ptr(col_begin), _ptr(col_end), _ptr(Wc), _ptr(lc), _ptr(plan), dst,
        len(dst_ptrs), int(row_offset), B, T, C.byref(sp), C.byref(mp), _dev(x), _stream(x)))


@torch.library.custom_op(f"{_NS}::mfcc_wave", mutates_args=(), device_types="cuda")
def mfcc_wave(x: Tensor, window: Tensor, H: Tensor, col_begin: Optional[Tensor], col_end: Optional[Tensor],
              W: Tensor, lifter: Tensor, frame_period: int, fft_length: int, center: bool, zmean: bool,
              pad_mode: int, eps: float, floor: float, gamma: float, out_format: int,
              plan: Optional[Tensor] = None) -> Tensor:
    dt = _native_dtype(x, window, H)
    n = num_frames(x.shape[-1], frame_period)
    y = torch.empty((*x.shape[:-1], n, _mfcc_dim(lifter.shape[-1] - 1, out_format)), device=x.device, dtype=dt)
    if y.numel():
        _mfcc_wave_call(x, window, H, col_begin, col_end, W, lifter, plan, [y.data_ptr()], 0, frame_period,
                        fft_length, center, zmean, pad_mode, eps, floor, gamma, out_format)
    return y


@mfcc_wave.register_fake
def _(x, window, H, col_begin, col_end, W, lifter, frame_period, fft_length, center, zmean, pad_mode, eps,
      floor, gamma, out_format, plan=None):
    return x.new_empty((*x.shape[:-1], num_frames(x.shape[-1], frame_period),
                        _mfcc_dim(lifter.shape[-1] - 1, out_format)), dtype=_native_dtype(x, window, H))


def mfcc_wave_gather(x: Tensor, out: Tensor, dst_ptrs, row_offset: int, window: Tensor, H: Tensor, col_begin, col_end,
                     W: Tensor, lifter: Tensor, frame_period: int, fft_length: int, center: bool, zmean: bool,
                     pad_mode: int, eps: float, floor: float, gamma: float, out_format: int, plan=None) -> Tensor:
    """Forward-only ``mfcc_wave`` whose rows land at row ``row_offset`` of ``out`` as mapped on every address in
    ``dst_ptrs`` (this rank's own tensor, its peers' mappings, or one multicast address): the all-gather of
    batch-sharded features inside the kernel's stores (``dsb200_mfcc_wave_ex``).  Synchronisation across ranks is
    the caller's (``distributed.FusedGatherMfcc``)."""
    if not 1 <= len(dst_ptrs) <= MAX_GATHER_PEERS:
        raise ValueError(f"between 1 and {MAX_GATHER_PEERS} destinations")
    if out.dtype != torch.float32 or not out.is_contiguous() or x.dim() != 2:
        raise ValueError("out must be a contiguous float32 tensor and x a [batch, T] tensor")
    n = num_frames(x.shape[-1], frame_period)
    D = _mfcc_dim(lifter.shape[-1] - 1, out_format)
    if out.shape[-1] != D or (row_offset + x.shape[0] * n) * D > out.numel():
        raise ValueError("out is too small for the rows of this call")
    _no_grad_check(x, window, H)
    if x.numel():
        _mfcc_wave_call(x, window, H, col_begin, col_end, W, lifter, plan, [int(p) for p in dst_ptrs], row_offset,
                        frame_period, fft_length, center, zmean, pad_mode, eps, floor, gamma, out_format)
    return out


# ------------------------------------------------------------------------- host-buffer pipeline
class HostStftPipeline:
    """Pinned host -> device -> fused STFT -> pinned host, chunked and overlapped (C ABI object)."""

    def __init__(self, window: Tensor, T: int, frame_period: int, fft_length: int, *, chunk_utterances: int = 32,
                 center: bool = True, zmean: bool = False, pad_mode: int = 0, eps: float = 1e-9,
                 relative_floor: Optional[float] = None, out_format: int = 3):
        if not window.is_cuda:
            raise RuntimeError("the window table must live on the CUDA device")
        self.window = window.contiguous()
        # the pipeline runs on its own non-blocking streams: make sure the table (possibly just produced on the
        # current torch stream) is complete before any of them can read it
        torch.cuda.current_stream(window.device).synchronize()
        self.is_f64 = window.dtype == torch.float64
        self.T, self.N, self.K = T, num_frames(T, frame_period), fft_length // 2 + 1
        self.out_format = out_format
        self._p = N.StftParams(_frame_params(window.shape[-1], frame_period, center, zmean, pad_mode),
                               _spec_params(fft_length, out_format, eps, relative_floor))
        self._h = C.c_void_p()
        N.check(N.load().dsb200_pipeline_create(C.byref(self._h), _dev(window), chunk_utterances, T,
                                                C.byref(self._p), int(self.is_f64)))

    def out_shape(self, batch: int):
        return (batch, self.N, self.K, 2) if self.out_format == 4 else (batch, self.N, self.K)

    def __call__(self, x_host: Tensor, y_host: Optional[Tensor] = None) -> Tensor:
        dt = torch.float64 if self.is_f64 else torch.float32
        if x_host.is_cuda or x_host.dtype != dt or not x_host.is_contiguous() or x_host.dim() != 2 or x_host.shape[1] != self.T:
            raise ValueError("x_host must be a contiguous CPU tensor of shape [batch, T] in the pipeline dtype")
        B = x_host.shape[0]
        if y_host is None:
            y_host = torch.empty(self.out_shape(B), dtype=dt, pin_memory=True)
        N.check(N.load().dsb200_pipeline_stft_host(self._h, C.c_void_p(x_host.data_ptr()), _ptr(self.window),
                                                   C.c_void_p(y_host.data_ptr()), B))
        return y_host

    def close(self):
        if self._h:
            N.load().dsb200_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------- autograd
# SURVEY.md section 8(f) rank 1.  frame, window, fftr, spec (numerator), stft, freqt, dct, acorr, levdur, lpc,
# fbank and mfcc (and the fused waveform pipelines) are differentiable through native adjoint kernels;
# mcep (the Newton solver) is forward-only and raises instead of silently dropping gradients.
def _prep_grad(g: Tensor, dtype: torch.dtype) -> Tensor:
    """Output gradient as a contiguous real tensor; complex gradients become interleaved (re, im) pairs."""
    if g.is_complex():
        g = torch.view_as_real(g.resolve_conj().contiguous())
    return _prep(g, dtype)


def _like_input(g: Tensor, x: Tensor) -> Tensor:
    return g.reshape(x.shape).to(x.dtype) if x.dtype.is_floating_point else None


@torch.library.custom_op(f"{_NS}::stft_backward", mutates_args=(), device_types="cuda")
def stft_backward(x: Tensor, window: Tensor, gy: Tensor, frame_period: int, fft_length: int, center: bool,
                  zmean: bool, pad_mode: int, eps: float, relative_floor: float, out_format: int,
                  need_gw: bool) -> tuple[Tensor, Tensor]:
    dt = _native_dtype(x, window)
    xc, wc, gc = _prep(x, dt), _prep(window, dt), _prep_grad(gy, dt)
    T = xc.shape[-1]
    B = xc.numel() // max(T, 1)
    gx = torch.empty_like(xc)
    gw = torch.empty(wc.shape if need_gw else (0,), device=x.device, dtype=dt)
    p = N.StftParams(_frame_params(wc.shape[-1], frame_period, center, zmean, pad_mode),
                     _spec_params(fft_length, out_format, eps, relative_floor))
    N.check(N.typed("dsb200_stft_backward", dt == torch.float64)(
        _ptr(xc), _ptr(wc), _ptr(gc), _ptr(gx), _ptr(gw) if need_gw else None, B, T, C.byref(p), _dev(x),
        _stream(x)))
    return gx, gw


@stft_backward.register_fake
def _(x, window, gy, frame_period, fft_length, center, zmean, pad_mode, eps, relative_floor, out_format, need_gw):
    dt = _native_dtype(x, window)
    return x.new_empty(x.shape, dtype=dt), x.new_empty(window.shape if need_gw else (0,), dtype=dt)


def _stft_setup(ctx, inputs, output):
    x, window, *rest = inputs
    ctx.save_for_backward(x, window)
    ctx.rest = rest


def _stft_bwd(ctx, g):
    x, window = ctx.saved_tensors
    need_gw = ctx.needs_input_grad[1]
    gx, gw = stft_backward(x, window, g, *ctx.rest, need_gw)
    return (_like_input(gx, x), gw.to(window.dtype) if need_gw else None) + (None,) * len(ctx.rest)


torch.library.register_autograd(f"{_NS}::stft", _stft_bwd, setup_context=_stft_setup)


@torch.library.custom_op(f"{_NS}::frame_backward", mutates_args=(), device_types="cuda")
def frame_backward(gy: Tensor, T: int, frame_period: int, center: bool, zmean: bool, pad_mode: int) -> Tensor:
    dt = _native_dtype(gy)
    gc = _prep(gy, dt)
    L = gc.shape[-1]
    lead = gc.shape[:-2]
    B = 1
    for s in lead:
        B *= s
    gx = torch.empty((*lead, T), device=gy.device, dtype=dt)
    p = _frame_params(L, frame_period, center, zmean, pad_mode)
    N.check(N.typed("dsb200_frame_backward", dt == torch.float64)(_ptr(gc), _ptr(gx), B, T, C.byref(p), _dev(gy),
                                                                  _stream(gy)))
    return gx


@frame_backward.register_fake
def _(gy, T, frame_period, center, zmean, pad_mode):
    return gy.new_empty((*gy.shape[:-2], T), dtype=_native_dtype(gy))


def _frame_setup(ctx, inputs, output):
    x, frame_length, frame_period, center, zmean, pad_mode = inputs
    ctx.x_meta = (x.shape, x.dtype)
    ctx.args = (frame_period, center, zmean, pad_mode)


def _frame_bwd(ctx, g):
    shape, dtype = ctx.x_meta
    gx = frame_backward(g, shape[-1], *ctx.args)
    return (gx.reshape(shape).to(dtype) if dtype.is_floating_point else None, None, None, None, None, None)


torch.library.register_autograd(f"{_NS}::frame", _frame_bwd, setup_context=_frame_setup)


@torch.library.custom_op(f"{_NS}::rfft_backward", mutates_args=(), device_types="cuda")
def rfft_backward(x: Tensor, gy: Tensor, fft_length: int, out_format: int) -> Tensor:
    dt = _native_dtype(x)
    xc, gc = _prep(x, dt), _prep_grad(gy, dt)
    Lin = xc.shape[-1]
    rows = xc.numel() // max(Lin, 1)
    gx = torch.empty_like(xc)
    N.check(N.typed("dsb200_rfft_backward", dt == torch.float64)(_ptr(xc), _ptr(gc), _ptr(gx), rows, Lin, fft_length,
                                                                 out_format, _dev(x), _stream(x)))
    return gx


@rfft_backward.register_fake
def _(x, gy, fft_length, out_format):
    return x.new_empty(x.shape, dtype=_native_dtype(x))


def _rfft_setup(ctx, inputs, output):
    x, fft_length, out_format = inputs
    ctx.save_for_backward(x)
    ctx.args = (fft_length, out_format)


def _rfft_bwd(ctx, g):
    (x,) = ctx.saved_tensors
    return _like_input(rfft_backward(x, g, *ctx.args), x), None, None


torch.library.register_autograd(f"{_NS}::rfft", _rfft_bwd, setup_context=_rfft_setup)


@torch.library.custom_op(f"{_NS}::spec_backward", mutates_args=(), device_types="cuda")
def spec_backward(b: Tensor, gy: Tensor, fft_length: int, eps: float, relative_floor: float, out_format: int) -> Tensor:
    dt = _native_dtype(b)
    bc, gc = _prep(b, dt), _prep_grad(gy, dt)
    Lb = bc.shape[-1]
    rows = bc.numel() // max(Lb, 1)
    gb = torch.empty_like(bc)
    p = _spec_params(fft_length, out_format, eps, relative_floor)
    N.check(N.typed("dsb200_spec_backward", dt == torch.float64)(_ptr(bc), Lb, _ptr(gc), _ptr(gb), rows, C.byref(p),
                                                                 _dev(b), _stream(b)))
    return gb


@spec_backward.register_fake
def _(b, gy, fft_length, eps, relative_floor, out_format):
    return b.new_empty(b.shape, dtype=_native_dtype(b))


def _spec_setup(ctx, inputs, output):
    b, a, *rest = inputs
    if a is not None and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
        # modules.Spectrum routes such calls to its differentiable composite; a direct call of the fused op cannot
        raise NotImplementedError("the fused pole-zero spectrum is forward-only: use modules.Spectrum / functional.spec, "
                                  "which differentiate b and a through the numerator kernel")
    ctx.save_for_backward(b)
    ctx.rest = rest


def _spec_bwd(ctx, g):
    (b,) = ctx.saved_tensors
    return (_like_input(spec_backward(b, g, *ctx.rest), b), None) + (None,) * len(ctx.rest)


torch.library.register_autograd(f"{_NS}::spec", _spec_bwd, setup_context=_spec_setup)


def _window_setup(ctx, inputs, output):
    x, w, out_length = inputs
    ctx.save_for_backward(x, w)


def _window_bwd(ctx, g):
    x, w = ctx.saved_tensors
    L1 = x.shape[-1]
    gs = g[..., :L1] if g.shape[-1] >= L1 else torch.nn.functional.pad(g, (0, L1 - g.shape[-1]))
    gx = window(gs, w, L1) if ctx.needs_input_grad[0] else None  # g * w with the same kernel
    gw = None
    if ctx.needs_input_grad[1]:
        gw = (gs.to(w.dtype) * x.to(w.dtype)).reshape(-1, L1).sum(0)
    return (_like_input(gx, x) if gx is not None else None), gw, None


torch.library.register_autograd(f"{_NS}::window", _window_bwd, setup_context=_window_setup)


def _rowmat_setup(ctx, inputs, output):
    x, W = inputs
    ctx.save_for_backward(x, W)


def _rowmat_bwd(ctx, g):
    x, W = ctx.saved_tensors
    gx = rowmat(g, W.t().contiguous()) if ctx.needs_input_grad[0] else None
    gW = None
    if ctx.needs_input_grad[1]:
        gW = x.reshape(-1, x.shape[-1]).to(W.dtype).t() @ g.reshape(-1, g.shape[-1]).to(W.dtype)
    return (_like_input(gx, x) if gx is not None else None), gW


torch.library.register_autograd(f"{_NS}::rowmat", _rowmat_bwd, setup_context=_rowmat_setup)


# ---- filter bank / MFCC (gradients with respect to the spectrum; the filter-bank matrix itself is not trained) ----
@torch.library.custom_op(f"{_NS}::fbank_backward", mutates_args=(), device_types="cuda")
def fbank_backward(x: Tensor, H: Tensor, col_begin: Optional[Tensor], col_end: Optional[Tensor], gy: Tensor,
                   gE: Optional[Tensor], floor: float, gamma: float, use_power: bool) -> Tensor:
    dt = _native_dtype(x, H)
    xc, Hc, gc = _prep(x, dt), _prep(H, dt), _prep(gy, dt)
    gEc = _prep(gE, dt) if gE is not None and gE.numel() else None
    K, Cn = Hc.shape
    rows = xc.numel() // max(K, 1)
    gx = torch.empty_like(xc)
    p = N.FbankParams(2 * (K - 1), Cn, int(use_power), int(gEc is not None), float(floor), float(gamma))
    N.check(N.typed("dsb200_fbank_backward", dt == torch.float64)(
        _ptr(xc), _ptr(Hc), _ptr(col_begin), _ptr(col_end), _ptr(gc), _ptr(gEc), _ptr(gx), rows, C.byref(p),
        _dev(x), _stream(x)))
    return gx


@fbank_backward.register_fake
def _(x, H, col_begin, col_end, gy, gE, floor, gamma, use_power):
    return x.new_empty(x.shape, dtype=_native_dtype(x, H))


def _fbank_weight_grad(P: Tensor, H: Tensor, gmel: Tensor, floor: float, gamma: float, use_power: bool) -> Tensor:
    """d/dH of the filter-bank outputs (learnable filter bank): amp^T @ (gmel * dy/dz), a plain dense GEMM."""
    amp = (P if use_power else torch.sqrt(P)).reshape(-1, P.shape[-1]).to(H.dtype)
    z = amp @ H
    dz = torch.where(z >= floor, (1.0 / z) if gamma == 0 else z.pow(gamma - 1.0), torch.zeros_like(z))
    return amp.t() @ (gmel.reshape(-1, H.shape[1]).to(H.dtype) * dz)


def _fbank_setup(ctx, inputs, output):
    x, H, cb, ce, floor, gamma, use_power, want_energy = inputs
    ctx.save_for_backward(x, H, cb, ce)
    ctx.rest = (floor, gamma, use_power)
    ctx.want_energy = want_energy


def _fbank_bwd(ctx, gy, gE):
    x, H, cb, ce = ctx.saved_tensors
    gx = fbank_backward(x, H, cb, ce, gy, gE if ctx.want_energy else None, *ctx.rest) if ctx.needs_input_grad[0] else None
    gH = _fbank_weight_grad(x, H, gy, *ctx.rest) if ctx.needs_input_grad[1] else None
    return (_like_input(gx, x) if gx is not None else None, gH) + (None,) * 6


torch.library.register_autograd(f"{_NS}::fbank", _fbank_bwd, setup_context=_fbank_setup)


def _mfcc_grad_to_spectrum(g: Tensor, P: Tensor, H, cb, ce, W, lifter, floor, gamma, out_format, want_gH=False):
    """Adjoint of lifter -> DCT -> log filter bank (mfcc.py:243-256): gradient of the packed output -> spectrum
    (and, for a learnable filter bank, -> H)."""
    M = lifter.shape[-1] - 1
    g = g.to(lifter.dtype)
    gcep = g.new_zeros((*g.shape[:-1], M + 1))
    gcep[..., 1:] = g[..., :M]
    if out_format in (2, 3):        # yc | ycE carry c0 right after the M cepstral coefficients
        gcep[..., 0] = g[..., M]
    gE = None
    if out_format == 1:
        gE = g[..., M].contiguous()
    elif out_format == 3:
        gE = g[..., M + 1].contiguous()
    gmel = rowmat(gcep * lifter, W[:, : M + 1].t().contiguous())
    gP = fbank_backward(P, H, cb, ce, gmel, gE, floor, gamma, False)
    return (gP, _fbank_weight_grad(P, H, gmel, floor, gamma, False)) if want_gH else (gP, None)


def _mfcc_setup(ctx, inputs, output):
    x, H, cb, ce, W, lifter, floor, gamma, out_format = inputs
    ctx.save_for_backward(x, H, cb, ce, W, lifter)
    ctx.rest = (floor, gamma, out_format)


def _mfcc_bwd(ctx, g):
    x, H, cb, ce, W, lifter = ctx.saved_tensors
    gx, gH = _mfcc_grad_to_spectrum(g, x, H, cb, ce, W, lifter, *ctx.rest, want_gH=ctx.needs_input_grad[1])
    return (_like_input(gx, x), gH) + (None,) * 7


torch.library.register_autograd(f"{_NS}::mfcc", _mfcc_bwd, setup_context=_mfcc_setup)


def _mfcc_wave_setup(ctx, inputs, output):
    (x, window, H, cb, ce, W, lifter, frame_period, fft_length, center, zmean, pad_mode, eps, floor, gamma,
     out_format, _plan) = inputs
    ctx.save_for_backward(x, window, H, cb, ce, W, lifter)
    ctx.stft_args = (frame_period, fft_length, center, zmean, pad_mode, eps, -1.0, 3)
    ctx.rest = (floor, gamma, out_format)


def _mfcc_wave_bwd(ctx, g):
    # The fused forward keeps nothing: recompute the power spectrum, pull the gradient back through the
    # filter bank, then through the STFT (all native kernels).
    x, window, H, cb, ce, W, lifter = ctx.saved_tensors
    with torch.no_grad():
        P = stft(x, window, *ctx.stft_args)
        gP, gH = _mfcc_grad_to_spectrum(g, P, H, cb, ce, W, lifter, *ctx.rest, want_gH=ctx.needs_input_grad[2])
        need_gw = ctx.needs_input_grad[1]
        gx, gw = stft_backward(x, window, gP, *ctx.stft_args, need_gw)
    return (_like_input(gx, x), gw.to(window.dtype) if need_gw else None, gH) + (None,) * 14


torch.library.register_autograd(f"{_NS}::mfcc_wave", _mfcc_wave_bwd, setup_context=_mfcc_wave_setup)


# ---- autocorrelation / Levinson-Durbin / LPC ----------------------------------------------------------------
@torch.library.custom_op(f"{_NS}::acorr_backward", mutates_args=(), device_types="cuda")
def acorr_backward(x: Tensor, gy: Tensor, acr_order: int, out_format: int) -> Tensor:
    dt = _native_dtype(x)
    xc, gc = _prep(x, dt), _prep(gy, dt)
    L = xc.shape[-1]
    rows = xc.numel() // max(L, 1)
    gx = torch.empty_like(xc)
    N.check(N.typed("dsb200_acorr_backward", dt == torch.float64)(_ptr(xc), _ptr(gc), _ptr(gx), rows, L, acr_order,
                                                                  out_format, _dev(x), _stream(x)))
    return gx


@acorr_backward.register_fake
def _(x, gy, acr_order, out_format):
    return x.new_empty(x.shape, dtype=_native_dtype(x))


@torch.library.custom_op(f"{_NS}::levdur_backward", mutates_args=(), device_types="cuda")
def levdur_backward(r: Tensor, ga: Tensor, eps: float) -> Tensor:
    dt = _native_dtype(r)
    rc, gc = _prep(r, dt), _prep(ga, dt)
    D = rc.shape[-1]
    rows = rc.numel() // max(D, 1)
    gr = torch.empty_like(rc)
    N.check(N.typed("dsb200_levdur_backward", dt == torch.float64)(_ptr(rc), _ptr(gc), _ptr(gr), rows, D - 1, eps,
                                                                   _dev(r), _stream(r)))
    return gr


@levdur_backward.register_fake
def _(r, ga, eps):
    return r.new_empty(r.shape, dtype=_native_dtype(r))


def _acorr_setup(ctx, inputs, output):
    x, acr_order, out_format = inputs
    ctx.save_for_backward(x)
    ctx.args = (acr_order, out_format)


def _acorr_bwd(ctx, g):
    (x,) = ctx.saved_tensors
    return _like_input(acorr_backward(x, g, *ctx.args), x), None, None


torch.library.register_autograd(f"{_NS}::acorr", _acorr_bwd, setup_context=_acorr_setup)


def _levdur_setup(ctx, inputs, output):
    r, eps = inputs
    ctx.save_for_backward(r)
    ctx.eps = eps


def _levdur_bwd(ctx, g):
    (r,) = ctx.saved_tensors
    return _like_input(levdur_backward(r, g, ctx.eps), r), None


torch.library.register_autograd(f"{_NS}::levdur", _levdur_bwd, setup_context=_levdur_setup)


def _lpc_setup(ctx, inputs, output):
    x, lpc_order, eps = inputs
    ctx.save_for_backward(x)
    ctx.args = (lpc_order, eps)


def _lpc_bwd(ctx, g):
    (x,) = ctx.saved_tensors
    M, eps = ctx.args
    with torch.no_grad():
        r = acorr(x, M, 0)
        gx = acorr_backward(x, levdur_backward(r, g, eps), M, 0)
    return _like_input(gx, x), None, None


torch.library.register_autograd(f"{_NS}::lpc", _lpc_bwd, setup_context=_lpc_setup)


def _lpc_wave_setup(ctx, inputs, output):
    x, window_t, frame_period, center, zmean, pad_mode, lpc_order, eps = inputs
    ctx.save_for_backward(x, window_t)
    ctx.frame_args = (frame_period, center, zmean, pad_mode)
    ctx.args = (lpc_order, eps)


def _lpc_wave_bwd(ctx, g):
    # Recompute the windowed frames (the fused forward never materialises them), then chain the adjoints.
    x, w = ctx.saved_tensors
    M, eps = ctx.args
    L = w.shape[-1]
    with torch.no_grad():
        fr = frame(x, L, *ctx.frame_args)
        fw = window(fr, w, L)
        r = acorr(fw, M, 0)
        gfw = acorr_backward(fw, levdur_backward(r, g, eps), M, 0)
        gfr = window(gfw, w, L)
        gx = frame_backward(gfr, x.shape[-1], *ctx.frame_args)
        gw = None
        if ctx.needs_input_grad[1]:
            gw = (gfw * fr).reshape(-1, L).sum(0).to(w.dtype)
    return (gx.reshape(x.shape).to(x.dtype), gw) + (None,) * 6


torch.library.register_autograd(f"{_NS}::lpc_wave", _lpc_wave_bwd, setup_context=_lpc_wave_setup)


# ------------------------------------------------------------------ inverse of the path (section 8f rank 2)
def _complex_rows(y: Tensor, dt: torch.dtype) -> Tensor:
    """Complex (or trailing-2 real) spectrum as contiguous interleaved (re, im) rows of dtype ``dt``."""
    if y.is_complex():
        y = torch.view_as_real(y.resolve_conj().contiguous())
    return _prep(y, dt)


def _complex_native_dtype(y: Tensor) -> torch.dtype:
    if y.is_complex():
        return torch.float64 if y.dtype == torch.complex128 else torch.float32
    return _native_dtype(y)


@torch.library.custom_op(f"{_NS}::ifftr", mutates_args=(), device_types="cuda")
def ifftr(y: Tensor, out_length: int) -> Tensor:
    """``y`` complex ``(..., K)`` -> real ``(..., out_length)``, ``fft_length = 2 (K - 1)``."""
    dt = _complex_native_dtype(y)
    yc = _complex_rows(y, dt)
    K = yc.shape[-2]
    rows = yc.numel() // max(2 * K, 1)
    x = torch.empty((*yc.shape[:-2], out_length), device=y.device, dtype=dt)
    N.check(N.typed("dsb200_ifftr", dt == torch.float64)(_ptr(yc), _ptr(x), rows, 2 * (K - 1), out_length, _dev(y),
                                                         _stream(y)))
    return x


@ifftr.register_fake
def _(y, out_length):
    return y.new_empty((*y.shape[:-1], out_length), dtype=_complex_native_dtype(y))


def unframe_length(n_frames: int, frame_length: int, frame_period: int, center: bool,
                   out_length: Optional[int]) -> int:
    """Length of the reference's ``x[..., s:e]`` slice of the folded signal (unframe.py:182-195)."""
    s = frame_length // 2 if center else 0
    avail = (n_frames - 1) * frame_period + frame_length - s
    if out_length is None:
        out_length = n_frames * frame_period if center else avail
    return max(0, min(out_length, avail))


@torch.library.custom_op(f"{_NS}::unframe", mutates_args=(), device_types="cuda")
def unframe(y: Tensor, window: Tensor, out_length: int, frame_period: int, center: bool) -> Tensor:
    dt = _native_dtype(y, window)
    yc, wc = _prep(y, dt), _prep(window, dt)
    Nf, L = yc.shape[-2], yc.shape[-1]
    B = yc.numel() // max(Nf * L, 1)
    out = torch.empty((*yc.shape[:-2], out_length), device=y.device, dtype=dt)
    if out_length > 0:
        N.check(N.typed("dsb200_unframe", dt == torch.float64)(_ptr(yc), _ptr(wc), _ptr(out), B, Nf, out_length, L,
                                                               frame_period, int(center), _dev(y), _stream(y)))
    return out


@unframe.register_fake
def _(y, window, out_length, frame_period, center):
    return y.new_empty((*y.shape[:-2], out_length), dtype=_native_dtype(y, window))


@torch.library.custom_op(f"{_NS}::istft", mutates_args=(), device_types="cuda")
def istft(y: Tensor, window: Tensor, out_length: int, frame_period: int, center: bool) -> Tensor:
    """``y`` complex ``(..., N, K)`` -> ``(..., out_length)``; the frames never reach HBM."""
    dt = _native_dtype(window)
    yc, wc = _complex_rows(y, dt), _prep(window, dt)
    Nf, K = yc.shape[-3], yc.shape[-2]
    B = yc.numel() // max(2 * Nf * K, 1)
    out = torch.empty((*yc.shape[:-3], out_length), device=y.device, dtype=dt)
    if out_length > 0:
        N.check(N.typed("dsb200_istft", dt == torch.float64)(_ptr(yc), _ptr(wc), _ptr(out), B, Nf, out_length,
                                                             wc.shape[-1], frame_period, 2 * (K - 1), int(center),
                                                             _dev(y), _stream(y)))
    return out


@istft.register_fake
def _(y, window, out_length, frame_period, center):
    return y.new_empty((*y.shape[:-2], out_length), dtype=_native_dtype(window))


# ---- gradients of the inverse path: the adjoints are the forward kernels of the analysis path -----------------
def _irfft_bin_weights(n: int, like: Tensor) -> Tensor:
    """c_k / n with c_k = 2 for the interior bins and 1 for DC / Nyquist (whose imaginary parts get no gradient)."""
    c = torch.full((n // 2 + 1,), 2.0 / n, device=like.device, dtype=like.dtype)
    c[0] = c[-1] = 1.0 / n
    return c


def _scale_half_spectrum(G: Tensor, n: int) -> Tensor:
    """interleaved (re, im) rows of rfft(g) -> gradient of irfft's complex input (torch convention dRe + i dIm)."""
    G = G * _irfft_bin_weights(n, G).unsqueeze(-1)
    G[..., 0, 1] = 0
    G[..., -1, 1] = 0
    return torch.view_as_complex(G.contiguous())


def _ifftr_setup(ctx, inputs, output):
    y, out_length = inputs
    ctx.n = 2 * (y.shape[-1] - 1)
    ctx.is_c64 = y.dtype == torch.complex64


def _ifftr_bwd(ctx, g):
    # x_j = (1/n) sum_k c_k Re(Y_k e^{+2 pi i jk/n})  =>  dL/dY = (c_k / n) rfft(g zero-padded to n)
    G = rfft(g, ctx.n, 0)
    out = _scale_half_spectrum(G, ctx.n)
    return out.to(torch.complex64 if ctx.is_c64 else torch.complex128), None


torch.library.register_autograd(f"{_NS}::ifftr", _ifftr_bwd, setup_context=_ifftr_setup)


def _ola_denominator(window: Tensor, n_frames: int, frame_period: int) -> Tensor:
    """sum_n w^2[q - n P] over the folded span (a batch-independent vector of (N-1) P + L values)."""
    L = window.shape[-1]
    w2 = (window.detach() * window.detach()).reshape(1, L, 1).expand(1, L, n_frames)
    span = (n_frames - 1) * frame_period + L
    return torch.nn.functional.fold(w2, (1, span), (1, L), stride=(1, frame_period)).reshape(span)


def _unframe_grad_signal(g: Tensor, window: Tensor, n_frames: int, frame_period: int, center: bool) -> Tensor:
    """g / (sum w^2 + 1e-16), laid out so that frame n of it starts where frame n was overlap-added."""
    L = window.shape[-1]
    s = L // 2 if center else 0
    den = _ola_denominator(window, n_frames, frame_period)
    T_out = g.shape[-1]
    u = g / (den[s:s + T_out] + 1e-16)
    need = (n_frames - 1) * frame_period + 1          # enough samples for the framing ops to produce N frames
    if T_out < need:
        u = torch.nn.functional.pad(u, (0, need - T_out))
    return u


def _synthesis_window_grad(u: Tensor, fr_u: Tensor, frames: Tensor, out: Tensor, w: Tensor, Nf: int, P: int,
                           center: bool) -> Tensor:
    """d/dw of out = fold(frames * w) / (fold(w^2) + 1e-16)  (unframe.py:198-204 of the reference).

    With u = g / den and t(n, j) the output sample that frame n's sample j lands on:
    dL/dw[j] = sum_n u[t(n, j)] (frames[n, j] - 2 w[j] out[t(n, j)]).  Both sums are framings of signals
    (native frame kernel); ``fr_u`` is the framed ``u`` the caller already has."""
    L = w.shape[-1]
    v = u.clone()
    To = out.shape[-1]
    v[..., :To] *= out
    v[..., To:] = 0
    fr_v = frame(v, L, P, center, False, 0)[..., :Nf, :]
    lead = tuple(range(fr_u.dim() - 1))
    return (fr_u * frames).sum(dim=lead) - 2.0 * w * fr_v.sum(dim=lead)


def _unframe_setup(ctx, inputs, output):
    y, window_t, out_length, frame_period, center = inputs
    ctx.save_for_backward(y, window_t, output)
    ctx.args = (frame_period, center)


def _unframe_bwd(ctx, g):
    y, w, out = ctx.saved_tensors
    P, center = ctx.args
    Nf, L = y.shape[-2], y.shape[-1]
    gw = None
    with torch.no_grad():
        u = _unframe_grad_signal(g, w, Nf, P, center)
        # d out[t] / d y[n, j] = w[j] / den(t) at t = n P + j - s: frame the scaled gradient, apply the window
        fr = frame(u, L, P, center, False, 0)[..., :Nf, :]
        gy = window(fr, w, L)
        if ctx.needs_input_grad[1]:   # learnable synthesis window
            gw = _synthesis_window_grad(u, fr, _prep(y, fr.dtype), out, w, Nf, P, center).to(w.dtype)
    return _like_input(gy, y), gw, None, None, None


torch.library.register_autograd(f"{_NS}::unframe", _unframe_bwd, setup_context=_unframe_setup)


def _istft_setup(ctx, inputs, output):
    y, window_t, out_length, frame_period, center = inputs
    if ctx.needs_input_grad[1]:
        ctx.save_for_backward(window_t, y, output)
    else:
        ctx.save_for_backward(window_t)
    ctx.meta = (y.shape[-2], 2 * (y.shape[-1] - 1), y.dtype == torch.complex64)
    ctx.args = (frame_period, center)


def _istft_bwd(ctx, g):
    # unframe's adjoint is window * frame(g / den), irfft's is (c_k / n) rfft: together, the complex STFT of
    # g / den -- the fused forward kernel -- scaled per bin.
    w = ctx.saved_tensors[0]
    Nf, n, is_c64 = ctx.meta
    P, center = ctx.args
    gw = None
    with torch.no_grad():
        u = _unframe_grad_signal(g, w, Nf, P, center)
        G = stft(u, w, P, n, center, False, 0, 0.0, -1.0, 4)[..., :Nf, :, :]
        out = _scale_half_spectrum(G, n)
        if ctx.needs_input_grad[1]:   # learnable synthesis window: the frames are ifftr(Y), recomputed here
            _, Y, x_out = ctx.saved_tensors
            L = w.shape[-1]
            fr_u = frame(u, L, P, center, False, 0)[..., :Nf, :]
            gw = _synthesis_window_grad(u, fr_u, ifftr(Y, L), x_out, w, Nf, P, center).to(w.dtype)
    return out.to(torch.complex64 if is_c64 else torch.complex128), gw, None, None, None


torch.library.register_autograd(f"{_NS}::istft", _istft_bwd, setup_context=_istft_setup)


# ------------------------------------------------------------------ delta features (section 8f rank 4)
def _delta_call(name: str, src: Tensor, window: Tensor, dim_out: int, D: int) -> Tensor:
    dt = _native_dtype(src, window)
    sc, wc = _prep(src, dt), _prep(window, dt)
    Tn = sc.shape[-2]
    B = sc.numel() // max(Tn * sc.shape[-1], 1)
    out = torch.empty((*sc.shape[:-1], dim_out), device=src.device, dtype=dt)
    N.check(N.typed(name, dt == torch.float64)(_ptr(sc), _ptr(wc), _ptr(out), B, Tn, D, wc.shape[0], wc.shape[1],
                                               _dev(src), _stream(src)))
    return out


@torch.library.custom_op(f"{_NS}::delta", mutates_args=(), device_types="cuda")
def delta(x: Tensor, window: Tensor) -> Tensor:
    """``x`` ``(..., T, D)``, ``window`` ``(H, W)`` -> ``(..., T, H D)`` (replicate padding over T)."""
    return _delta_call("dsb200_delta", x, window, window.shape[0] * x.shape[-1], x.shape[-1])


@delta.register_fake
def _(x, window):
    return x.new_empty((*x.shape[:-1], window.shape[0] * x.shape[-1]), dtype=_native_dtype(x, window))


@torch.library.custom_op(f"{_NS}::delta_backward", mutates_args=(), device_types="cuda")
def delta_backward(gy: Tensor, window: Tensor) -> Tensor:
    D = gy.shape[-1] // window.shape[0]
    return _delta_call("dsb200_delta_backward", gy, window, D, D)


@delta_backward.register_fake
def _(gy, window):
    return gy.new_empty((*gy.shape[:-1], gy.shape[-1] // window.shape[0]), dtype=_native_dtype(gy, window))


def _delta_setup(ctx, inputs, output):
    x, window_t = inputs
    ctx.save_for_backward(x, window_t)


def _delta_bwd(ctx, g):
    x, w = ctx.saved_tensors
    if ctx.needs_input_grad[1]:
        raise NotImplementedError("gradients with respect to the regression window are not implemented")
    return _like_input(delta_backward(g, w), x), None


torch.library.register_autograd(f"{_NS}::delta", _delta_bwd, setup_context=_delta_setup)


# ------------------------------------------------------------------ per-row converters (section 8f rank 4)
CONV_LPC2PAR, CONV_PAR2LPC, CONV_GNORM, CONV_IGNORM, CONV_NORM0 = range(5)


@torch.library.custom_op(f"{_NS}::rowconv", mutates_args=(), device_types="cuda")
def rowconv(x: Tensor, op: int, param: float) -> Tensor:
    """``(..., D) -> (..., D)``: lpc2par / par2lpc / gnorm / ignorm / norm0 on every row (``dsb200_rowconv``)."""
    dt = _native_dtype(x)
    xc = _prep(x, dt)
    D = xc.shape[-1]
    rows = xc.numel() // max(D, 1)
    y = torch.empty_like(xc)
    N.check(N.typed("dsb200_rowconv", dt == torch.float64)(_ptr(xc), _ptr(y), rows, D, int(op), float(param), _dev(x),
                                                           _stream(x)))
    return y


@rowconv.register_fake
def _(x, op, param):
    return x.new_empty(x.shape, dtype=_native_dtype(x))


def rowconv_composite(x: Tensor, op: int, g: float) -> Tensor:
    """The same five converters as differentiable torch expressions on the device: used by the backward of
    ``rowconv`` only (recompute + torch.autograd), never by a forward call."""
    M = x.shape[-1] - 1
    K, a = torch.split(x, [1, M], dim=-1)
    if op == CONV_LPC2PAR:          # lpc2par.py:104-120
        ks = []
        a = a * g
        for m in reversed(range(M)):
            km = a[..., m:m + 1]
            ks.append(km)
            if m == 0:
                break
            k = a[..., :-1]
            a = (k - km * k.flip(-1)) / (1 - km * km)
        ks.append(K)
        return torch.cat(ks[::-1], dim=-1)
    if op == CONV_PAR2LPC:          # par2lpc.py:100-107
        cols = [x[..., i:i + 1] / g for i in range(min(2, M + 1))]
        for m in range(2, M + 1):
            km = x[..., m:m + 1]
            am = torch.cat(cols[1:m], dim=-1)
            am = am + km * am.flip(-1)
            cols = cols[:1] + list(torch.split(am, 1, dim=-1)) + [km / g]
        return torch.cat(cols, dim=-1)
    if op == CONV_GNORM:            # gnorm.py:101-112
        if g == 0:
            return torch.cat((torch.exp(K), a), dim=-1)
        z = 1 + g * K
        return torch.cat((torch.pow(z, 1 / g), a / z), dim=-1)
    if op == CONV_IGNORM:           # ignorm.py:98-109
        if g == 0:
            return torch.cat((torch.log(K), a), dim=-1)
        z = torch.pow(K, g)
        return torch.cat(((z - 1) / g, a * z), dim=-1)
    b0 = torch.reciprocal(K)        # norm0.py:88-94
    return torch.cat((b0, a * b0), dim=-1)


def _rowconv_setup(ctx, inputs, output):
    x, op, param = inputs
    ctx.save_for_backward(x)
    ctx.op, ctx.param = op, param


def _rowconv_bwd(ctx, g):
    (x,) = ctx.saved_tensors
    with torch.enable_grad():
        xd = x.detach().to(_native_dtype(x)).requires_grad_(True)
        y = rowconv_composite(xd, ctx.op, ctx.param)
        (gx,) = torch.autograd.grad(y, xd, g.to(y.dtype))
    return _like_input(gx, x), None, None


torch.library.register_autograd(f"{_NS}::rowconv", _rowconv_bwd, setup_context=_rowconv_setup)


# ------------------------------------------------------------------ Toeplitz-plus-Hankel solve (mgcep Newton step)
@torch.library.custom_op(f"{_NS}::thsolve", mutates_args=(), device_types="cuda")
def thsolve(t: Tensor, h: Tensor, r: Tensor) -> Tensor:
    """``(Toeplitz(t) + Hankel(h)) x = r`` per row: ``t (..., M)``, ``h (..., 2M-1)``, ``r (..., M) -> x (..., M)``."""
    dt = _native_dtype(t, h, r)
    tc, hc, rc = _prep(t, dt), _prep(h, dt), _prep(r, dt)
    M = tc.shape[-1]
    if hc.shape[-1] != 2 * M - 1 or rc.shape[-1] != M:
        raise ValueError("thsolve: h must have 2 M - 1 and r must have M elements per row.")
    rows = tc.numel() // max(M, 1)
    x = torch.empty_like(rc)
    N.check(N.typed("dsb200_thsolve", dt == torch.float64)(_ptr(tc), _ptr(hc), _ptr(rc), _ptr(x), rows, M, _dev(t),
                                                           _stream(t)))
    return x


@thsolve.register_fake
def _(t, h, r):
    return r.new_empty(r.shape, dtype=_native_dtype(t, h, r))


def _thsolve_setup(ctx, inputs, output):
    t, h, r = inputs
    ctx.save_for_backward(t, h, output)


def _thsolve_bwd(ctx, g):
    # A x = r with A = T(t) + H(h) symmetric:  lam = A^-1 g (the same solver);  dr = lam;  dA = -lam x^T;
    # dt_d = sum_{|i-j|=d} dA_ij,  dh_s = sum_{i+j=s} dA_ij
    t, h, x = ctx.saved_tensors
    lam = thsolve(t, h, g.to(x.dtype))
    dA = -lam.unsqueeze(-1) * x.unsqueeze(-2)
    M = x.shape[-1]
    i = torch.arange(M, device=x.device)
    dist = (i[:, None] - i[None, :]).abs().reshape(-1)
    summ = (i[:, None] + i[None, :]).reshape(-1)
    flat = dA.reshape(*dA.shape[:-2], M * M)
    dt = torch.zeros_like(t, dtype=x.dtype).index_add_(-1, dist, flat)
    dh = torch.zeros_like(h, dtype=x.dtype).index_add_(-1, summ, flat)
    return _like_input(dt, t), _like_input(dh, h), _like_input(lam, x)


torch.library.register_autograd(f"{_NS}::thsolve", _thsolve_bwd, setup_context=_thsolve_setup)


# ------------------------------------------------------------------ mcep gradients (recompute + autograd)
def mcep_composite(x: Tensor, P0: Tensor, G: Tensor, Hm: Tensor, alpha_vector: Tensor, n_iter: int) -> Tensor:
    """The Newton iteration of the fused mcep kernels, step by step on DIFFERENTIABLE kernels of this package
    (``rowmat`` with the folded tables, ``thsolve``) and device-side elementwise ops -- used by the backward of
    ``mcep`` only (the forward is one fused kernel that keeps nothing for autograd)."""
    D = alpha_vector.shape[-1]
    lx = torch.log(x)
    mc = rowmat(lx, P0)
    for _ in range(n_iter):
        e = torch.exp(lx - 2 * rowmat(mc, G))
        rt = rowmat(e, Hm)
        mc = mc + thsolve(rt[..., :D], rt, rt[..., :D] - alpha_vector)
    return mc


def _mcep_setup(ctx, inputs, output):
    x, P0, G, Hm, alpha_vector, n_iter = inputs
    ctx.save_for_backward(x, P0, G, Hm, alpha_vector)
    ctx.n_iter = n_iter


def _mcep_bwd(ctx, g):
    x, P0, G, Hm, av = ctx.saved_tensors
    if any(ctx.needs_input_grad[1:5]):
        raise NotImplementedError("gradients with respect to the mel-cepstral analysis tables are not implemented")
    dt = _native_dtype(x, av)
    K = x.shape[-1]
    xf = x.detach().to(dt).reshape(-1, K)
    gf = g.to(dt).reshape(-1, g.shape[-1])
    tabs = [t.detach().to(dt) for t in (P0, G, Hm, av)]
    out = torch.empty_like(xf)
    chunk = 32768   # rows per recompute: bounds the autograd tape (~ n_iter * 350 floats per row)
    for lo in range(0, xf.shape[0], chunk):
        with torch.enable_grad():
            xd = xf[lo:lo + chunk].clone().requires_grad_(True)
            y = mcep_composite(xd, *tabs, ctx.n_iter)
            (out[lo:lo + chunk],) = torch.autograd.grad(y, xd, gf[lo:lo + chunk])
    return _like_input(out, x), None, None, None, None, None


torch.library.register_autograd(f"{_NS}::mcep", _mcep_bwd, setup_context=_mcep_setup)


# ------------------------------------------------------------------ lpc2lsp (section 8f rank 4)
@torch.library.custom_op(f"{_NS}::lpc2lsp", mutates_args=(), device_types="cuda")
def lpc2lsp(a: Tensor, log_gain: bool, scale: float) -> Tensor:
    """``[K, a_1..a_M] -> [K or log K, scale * w_1..w_M]`` (ascending line spectral frequencies)."""
    dt = _native_dtype(a)
    ac = _prep(a, dt)
    D = ac.shape[-1]
    rows = ac.numel() // max(D, 1)
    w = torch.empty_like(ac)
    N.check(N.typed("dsb200_lpc2lsp", dt == torch.float64)(_ptr(ac), _ptr(w), rows, D - 1, int(log_gain), float(scale),
                                                           _dev(a), _stream(a)))
    return w


@lpc2lsp.register_fake
def _(a, log_gain, scale):
    return a.new_empty(a.shape, dtype=_native_dtype(a))


def _lpc2lsp_setup(ctx, inputs, output):
    a, log_gain, scale = inputs
    ctx.save_for_backward(a, output)
    ctx.log_gain, ctx.scale = log_gain, scale


def _lpc2lsp_bwd(ctx, g):
    # Implicit differentiation.  With h = (M + 1) / 2 and a_0 = 1 the zeros of Q are the zeros of
    # R(w) = sum_k a_k cos((h - k) w) and the zeros of P those of I(w) = sum_k a_k sin((h - k) w)
    # (e^{j h w} A(e^{jw}) = R + j I):  dw/da_k = -cos((h - k) w) / R'(w)  resp.  -sin((h - k) w) / I'(w).
    a, out = ctx.saved_tensors
    dt = out.dtype
    M = a.shape[-1] - 1
    K = a[..., :1].to(dt)
    gK = g[..., :1].to(dt) / K if ctx.log_gain else g[..., :1].to(dt)
    if M == 0:
        return _like_input(gK, a), None, None
    w = (out[..., 1:] / ctx.scale).unsqueeze(-1)                       # (..., M, 1) radians
    k = torch.arange(M + 1, device=a.device, dtype=dt)
    hk = (M + 1) / 2 - k                                               # (M + 1,)
    coef = torch.cat((torch.ones_like(K), a[..., 1:].to(dt)), dim=-1).unsqueeze(-2)   # (..., 1, M + 1)
    c, s = torch.cos(hk * w), torch.sin(hk * w)                        # (..., M, M + 1)
    R, I = (coef * c).sum(-1), (coef * s).sum(-1)                      # (..., M)
    is_q = R.abs() < I.abs()
    num = torch.where(is_q.unsqueeze(-1), c, s)
    den = torch.where(is_q, -(coef * hk * s).sum(-1), (coef * hk * c).sum(-1))
    dw_da = -num / den.unsqueeze(-1)                                   # (..., M zeros, M + 1 coefficients)
    ga = (g[..., 1:].to(dt).unsqueeze(-1) * ctx.scale * dw_da).sum(-2)[..., 1:]
    return _like_input(torch.cat((gK, ga), dim=-1), a), None, None


torch.library.register_autograd(f"{_NS}::lpc2lsp", _lpc2lsp_bwd, setup_context=_lpc2lsp_setup)


# ------------------------------------------------------------------ gc2gc (the gamma conversion inside mgc2mgc)
def gc2gc_composite(c1: Tensor, out_order: int, in_gamma: float, out_gamma: float, n_fft: int) -> Tensor:
    """mgc2mgc.py:327-364 step by step on the differentiable FFT kernels (``rfft`` / ``ifftr``) and device-side
    elementwise ops: the backward of ``gc2gc`` (and its cross-check in the tests); even ``n_fft`` only."""
    c01 = torch.cat((torch.zeros_like(c1[..., :1]), c1[..., 1:]), dim=-1)
    C1 = torch.view_as_complex(rfft(c01, n_fft, 0))     # the other half of fft(c01) is its mirror image
    if in_gamma == 0:
        sC1 = torch.polar(torch.exp(C1.real), C1.imag)
    else:
        C1 = torch.complex(C1.real * in_gamma + 1, C1.imag * in_gamma)
        sC1 = torch.polar(C1.abs() ** (1 / in_gamma), C1.angle() / in_gamma)
    if out_gamma == 0:
        C2 = torch.log(sC1.abs())
    else:
        C2 = ((sC1.abs() ** out_gamma) * torch.cos(sC1.angle() * out_gamma) - 1) / out_gamma
    c02 = ifftr(torch.complex(C2, torch.zeros_like(C2)), n_fft)[..., : out_order + 1]   # C2 is real and even
    return torch.cat((c1[..., :1], 2 * c02[..., 1:]), dim=-1)


@torch.library.custom_op(f"{_NS}::gc2gc", mutates_args=(), device_types="cuda")
def gc2gc(c1: Tensor, out_order: int, in_gamma: float, out_gamma: float, n_fft: int) -> Tensor:
    dt = _native_dtype(c1)
    cc = _prep(c1, dt)
    D1 = cc.shape[-1]
    rows = cc.numel() // max(D1, 1)
    c2 = torch.empty((*cc.shape[:-1], out_order + 1), device=c1.device, dtype=dt)
    N.check(N.typed("dsb200_gc2gc", dt == torch.float64)(_ptr(cc), _ptr(c2), rows, D1 - 1, out_order, float(in_gamma),
                                                         float(out_gamma), n_fft, _dev(c1), _stream(c1)))
    return c2


@gc2gc.register_fake
def _(c1, out_order, in_gamma, out_gamma, n_fft):
    return c1.new_empty((*c1.shape[:-1], out_order + 1), dtype=_native_dtype(c1))


def _gc2gc_setup(ctx, inputs, output):
    c1, out_order, in_gamma, out_gamma, n_fft = inputs
    ctx.save_for_backward(c1)
    ctx.args = (out_order, in_gamma, out_gamma, n_fft)


def _gc2gc_bwd(ctx, g):
    (c1,) = ctx.saved_tensors
    if ctx.args[3] % 2:
        raise NotImplementedError("gradients of gc2gc need an even n_fft (they run on the real-FFT kernels)")
    with torch.enable_grad():
        cd = c1.detach().to(_native_dtype(c1)).requires_grad_(True)
        y = gc2gc_composite(cd, *ctx.args)
        (gc,) = torch.autograd.grad(y, cd, g.to(y.dtype))
    return _like_input(gc, c1), None, None, None, None


torch.library.register_autograd(f"{_NS}::gc2gc", _gc2gc_bwd, setup_context=_gc2gc_setup)
