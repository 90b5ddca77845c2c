"""ctypes binding of the C ABI declared in ``include/diffsptk_b200.h``.

The shared library is the product: there is no Python/CPU fallback.  If the
library is missing, ``load()`` raises with the build command instead of
silently degrading.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", os.environ.get("DSB200_LIB_NAME", "libdiffsptk_b200.so"))

OK, E_BAD_PARAM, E_UNSUPPORTED, E_ALIGN, E_CUDA = 0, -1, -2, -3, -4

PAD_MODES = {"constant": 0, "reflect": 1, "replicate": 2, "circular": 3}


class FrameParams(C.Structure):
    _fields_ = [("frame_length", C.c_int32), ("frame_period", C.c_int32), ("center", C.c_int32),
                ("zmean", C.c_int32), ("pad_mode", C.c_int32)]


class SpecParams(C.Structure):
    _fields_ = [("fft_length", C.c_int32), ("out_format", C.c_int32), ("has_relative_floor", C.c_int32),
                ("reserved", C.c_int32), ("eps", C.c_double), ("relative_floor", C.c_double)]


class StftParams(C.Structure):
    _fields_ = [("frame", FrameParams), ("spec", SpecParams)]


class FbankParams(C.Structure):
    _fields_ = [("fft_length", C.c_int32), ("n_channel", C.c_int32), ("use_power", C.c_int32),
                ("want_energy", C.c_int32), ("floor", C.c_double), ("gamma", C.c_double)]


class MfccParams(C.Structure):
    _fields_ = [("fbank", FbankParams), ("mfcc_order", C.c_int32), ("out_format", C.c_int32)]


class McepParams(C.Structure):
    _fields_ = [("fft_length", C.c_int32), ("cep_order", C.c_int32), ("n_iter", C.c_int32),
                ("reserved", C.c_int32)]


_P, _I32, _I64, _D, _INT = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_int

# name -> argument types (without the _f32/_f64 suffix); every typed entry point returns int
_TYPED = {
    "dsb200_frame": [_P, _P, _I64, _I64, C.POINTER(FrameParams), _INT, _P],
    "dsb200_window": [_P, _P, _P, _I64, _I32, _I32, _INT, _P],
    "dsb200_rfft": [_P, _P, _I64, _I32, _I32, _I32, _INT, _P],
    "dsb200_spec": [_P, _I32, _P, _I32, _P, _I64, C.POINTER(SpecParams), _INT, _P],
    "dsb200_stft": [_P, _P, _P, _I64, _I64, C.POINTER(StftParams), _INT, _P],
    "dsb200_acorr": [_P, _P, _I64, _I32, _I32, _I32, _INT, _P],
    "dsb200_levdur": [_P, _P, _I64, _I32, _D, _INT, _P],
    "dsb200_lpc": [_P, _P, _I64, _I32, _I32, _D, _INT, _P],
    "dsb200_lpc_wave": [_P, _P, _P, _I64, _I64, C.POINTER(FrameParams), _I32, _D, _INT, _P],
    "dsb200_rowmat": [_P, _P, _P, _I64, _I32, _I32, _INT, _P],
    "dsb200_mcep": [_P, _P, _I64, C.POINTER(McepParams), _P, _P, _P, _P, _INT, _P],
    "dsb200_fbank": [_P, _P, _P, _P, _P, _P, _I64, C.POINTER(FbankParams), _INT, _P],
    "dsb200_mfcc": [_P, _P, _P, _P, _P, _P, _P, _I64, C.POINTER(MfccParams), _INT, _P],
    "dsb200_stft_backward": [_P, _P, _P, _P, _P, _I64, _I64, C.POINTER(StftParams), _INT, _P],
    "dsb200_rfft_backward": [_P, _P, _P, _I64, _I32, _I32, _I32, _INT, _P],
    "dsb200_spec_backward": [_P, _I32, _P, _P, _I64, C.POINTER(SpecParams), _INT, _P],
    "dsb200_frame_backward": [_P, _P, _I64, _I64, C.POINTER(FrameParams), _INT, _P],
    "dsb200_ifftr": [_P, _P, _I64, _I32, _I32, _INT, _P],
    "dsb200_unframe": [_P, _P, _P, _I64, _I64, _I64, _I32, _I32, _I32, _INT, _P],
    "dsb200_istft": [_P, _P, _P, _I64, _I64, _I64, _I32, _I32, _I32, _I32, _INT, _P],
    "dsb200_delta": [_P, _P, _P, _I64, _I64, _I32, _I32, _I32, _INT, _P],
    "dsb200_rowconv": [_P, _P, _I64, _I32, _I32, _D, _INT, _P],
    "dsb200_thsolve": [_P, _P, _P, _P, _I64, _I32, _INT, _P],
    "dsb200_lpc2lsp": [_P, _P, _I64, _I32, _I32, _D, _INT, _P],
    "dsb200_gc2gc": [_P, _P, _I64, _I32, _I32, _D, _D, _I32, _INT, _P],
    "dsb200_delta_backward": [_P, _P, _P, _I64, _I64, _I32, _I32, _I32, _INT, _P],
    "dsb200_fbank_backward": [_P, _P, _P, _P, _P, _P, _P, _I64, C.POINTER(FbankParams), _INT, _P],
    "dsb200_acorr_backward": [_P, _P, _P, _I64, _I32, _I32, _I32, _INT, _P],
    "dsb200_levdur_backward": [_P, _P, _P, _I64, _I32, _D, _INT, _P],
    "dsb200_mfcc_wave": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, C.POINTER(StftParams),
                         C.POINTER(MfccParams), _INT, _P],
    "dsb200_mfcc_wave_ex": [_P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_void_p), _I32, _I64, _I64, _I64,
                            C.POINTER(StftParams), C.POINTER(MfccParams), _INT, _P],
}

_PLAIN = {
    "dsb200_version": (C.c_int, []),
    "dsb200_last_error": (C.c_char_p, []),
    "dsb200_launch_count": (C.c_int64, []),
    "dsb200_last_kernel": (C.c_char_p, []),
    "dsb200_set_sm_margin": (C.c_int, [_I32]),
    "dsb200_set_knob": (C.c_int, [C.c_char_p, _I32]),
    "dsb200_clear_knobs": (C.c_int, []),
    "dsb200_num_frames": (C.c_int64, [_I64, _I32]),
    "dsb200_pipeline_create": (C.c_int, [C.POINTER(C.c_void_p), _INT, _I64, _I64, C.POINTER(StftParams), _INT]),
    "dsb200_pipeline_stft_host": (C.c_int, [_P, _P, _P, _P, _I64]),
    "dsb200_pipeline_destroy": (C.c_int, [_P]),
    "dsb200_mfcc_plan_ints": (C.c_int32, [_I32]),
    "dsb200_mfcc_plan_build": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32), _I32, _I32,
                                         C.POINTER(C.c_int32)]),
}


def exported_symbols():
    """Every symbol ``include/diffsptk_b200.h`` declares."""
    names = list(_PLAIN)
    for base in _TYPED:
        names += [base + "_f32", base + "_f64"]
    return names


_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the shared library (once).  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"diffsptk_b200: native library not found at {LIB_PATH}. Build it with "
                "`python -m diffsptk_b200.build` (needs nvcc); there is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PLAIN.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        for base, args in _TYPED.items():
            for suf in ("_f32", "_f64"):
                fn = getattr(lib, base + suf)
                fn.restype, fn.argtypes = C.c_int, args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    """Translate a dsb200_status into the exception the reference would raise."""
    if rc == OK:
        return
    msg = load().dsb200_last_error().decode("utf-8", "replace")
    if rc == E_BAD_PARAM:
        raise ValueError(msg)
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"diffsptk_b200 native error {rc}: {msg}")


def typed(base: str, is_f64: bool):
    return getattr(load(), base + ("_f64" if is_f64 else "_f32"))


def launch_count() -> int:
    return int(load().dsb200_launch_count())


def set_sm_margin(n_sms: int) -> int:
    """Leave ``n_sms`` SMs free of this library's persistent kernels (for a collective that overlaps them); returns
    the previous margin."""
    rc = int(load().dsb200_set_sm_margin(int(n_sms)))
    if rc < 0:
        check(rc)
    return rc


def set_knob(name: str, value: int) -> None:
    """Set a tuning knob of the kernels (README.md's table, without the ``DSB200_`` prefix); wins over the environment."""
    check(load().dsb200_set_knob(name.encode(), int(value)))


def clear_knobs() -> None:
    """Forget every knob set through :func:`set_knob`."""
    check(load().dsb200_clear_knobs())


def last_kernel() -> str:
    """Name of the kernel this thread launched last (which path served the call)."""
    return load().dsb200_last_kernel().decode()
