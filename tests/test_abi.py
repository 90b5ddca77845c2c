"""The C-ABI library builds for sm_100a, loads, and exports every symbol the header declares."""

import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "diffsptk_b200.h")).read()
    names = set(re.findall(r"DSB200_API\s+[\w\s\*]+?\b(dsb200_\w+)\s*\(", txt))
    for base in re.findall(r"DSB200_DECL2\((dsb200_\w+),", txt):
        names |= {base + "_f32", base + "_f64"}
    return names


def test_library_exports_header_symbols(native_lib):
    from diffsptk_b200 import _native
    hdr = header_symbols()
    assert len(hdr) >= 30
    assert hdr == set(_native.exported_symbols()), "ctypes binding and header disagree"
    for name in sorted(hdr):
        assert hasattr(native_lib, name), f"{name} is declared in the header but not exported"


def test_version_and_frame_count(native_lib):
    assert native_lib.dsb200_version() == 100
    f = native_lib.dsb200_num_frames
    # N = (T-1)//P + 1
    assert [f(t, 80) for t in (1, 79, 80, 81, 399, 400, 401, 160000, 160001)] == [1, 1, 1, 2, 5, 5, 6, 2000, 2001]
    assert f(0, 80) == 0


def test_parameter_errors_without_gpu(native_lib):
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only box."""
    import ctypes as C
    from diffsptk_b200 import _native as N
    p = N.FrameParams(0, 80, 1, 0, 0)
    rc = native_lib.dsb200_frame_f32(None, None, 1, 100, C.byref(p), 0, None)
    assert rc == N.E_BAD_PARAM and b"frame_length must be positive" in native_lib.dsb200_last_error()
    with pytest.raises(ValueError):
        N.check(rc)
    rc = native_lib.dsb200_rfft_f32(None, None, 1, 8, 7, 0, 0, None)
    assert rc == N.E_BAD_PARAM and b"even" in native_lib.dsb200_last_error()
    rc = native_lib.dsb200_acorr_f64(None, None, 1, 10, 10, 0, 0, None)
    assert rc == N.E_BAD_PARAM
    # zero rows is a no-op that needs no device
    assert native_lib.dsb200_window_f32(None, None, None, 0, 8, 8, 0, None) == 0


def test_sm_margin_setter(native_lib):
    """Process-wide setter (no GPU needed): returns the previous value, rejects negatives."""
    from diffsptk_b200 import _native as N
    assert N.set_sm_margin(8) == 0
    assert N.set_sm_margin(0) == 8
    assert native_lib.dsb200_set_sm_margin(-1) == N.E_BAD_PARAM
    assert N.set_sm_margin(0) == 0


def test_knob_setter(native_lib):
    """Tuning knobs are host-side state: set / clear / reject an empty name without any CUDA call."""
    from diffsptk_b200 import _native as N
    N.set_knob("LPC_STAGGER", 12000)
    N.set_knob("LPC_STAGGER", 0)
    N.clear_knobs()
    assert native_lib.dsb200_set_knob(b"", 1) == N.E_BAD_PARAM
    assert b"knob" in native_lib.dsb200_last_error()
    assert native_lib.dsb200_set_knob(None, 1) == N.E_BAD_PARAM


def test_sass_is_sm100a(native_lib):
    from diffsptk_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without the built .so the package must raise, not degrade."""
    from diffsptk_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="native library not found"):
        _native.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under diffsptk_b200/ may reference it."""
    pkg = os.path.join(ROOT, "diffsptk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "np_oracle" not in txt, f
