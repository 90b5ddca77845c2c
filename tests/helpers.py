"""Shared test helpers: golden-case loading, tolerances, op dispatch."""

from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

# The reference's own tolerances (tests/utils.py:66-72 of the reference).
TOL = {"f32": dict(rtol=1e-4, atol=1e-6), "f64": dict(rtol=1e-5, atol=1e-8)}

_manifest = None
_store = {}


def manifest():
    global _manifest
    if _manifest is None:
        with open(os.path.join(GOLDEN, "manifest.json")) as f:
            _manifest = json.load(f)
    return _manifest


def store(prec):
    if prec not in _store:
        _store[prec] = np.load(os.path.join(GOLDEN, f"golden_{prec}.npz"))
    return _store[prec]


def case_names(ops=None):
    m = manifest()
    return sorted(n for n, c in m.items() if ops is None or c["op"] in ops)


def load_case(name, prec):
    """-> (op, params, inputs(list, None for absent), outputs(list))"""
    c = manifest()[name]
    st = store(prec)
    ins = [None if i in c["none_in"] else st[f"{name}/in{i}"] for i in range(c["n_in"])]
    outs = [st[f"{name}/out{i}"] for i in range(c["n_out"])]
    return c["op"], dict(c["params"]), ins, outs


def assert_close(got, want, prec, what="", scale_atol=False, rtol_mul=1.0, atol_mul=1.0):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if np.iscomplexobj(want) or np.iscomplexobj(got):
        got = np.stack([got.real, got.imag], -1)
        want = np.stack([want.real, want.imag], -1)
    tol = TOL[prec]
    rtol, atol = tol["rtol"] * rtol_mul, tol["atol"] * atol_mul
    if scale_atol and want.size:
        finite = want[np.isfinite(want)]
        if finite.size:
            atol = atol * max(1.0, float(np.max(np.abs(finite))))
    ok = np.isclose(got, want.astype(got.dtype if got.dtype.kind == "f" else want.dtype), rtol=rtol, atol=atol,
                    equal_nan=True)
    if not ok.all():
        bad = np.argwhere(~ok)
        i = tuple(bad[0])
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        raise AssertionError(
            f"{what}: {len(bad)}/{ok.size} elements outside rtol={rtol:g} atol={atol:g}; first at {i}: "
            f"got {got[i]!r} want {want[i]!r}; max abs err {np.nanmax(err):.3e}")


def run_oracle(op, params, inputs):
    from oracle import np_oracle as O
    return getattr(O, op)(*inputs, **params)


def assert_close_conditioned(got, ref32, ref64, what="", factor=4.0):
    """float32 criterion for conditioning-dominated ops (levdur / lpc / mcep).

    The reference's own float32 output deviates from its float64 output by an amount set by the
    conditioning of each frame's linear system.  A float32 implementation is accepted when its
    distance to the float64 reference is within the standard float32 tolerance plus ``factor``
    times the reference's own float32 error on the same row.
    """
    got = np.asarray(got, dtype=np.float64)
    ref32 = np.asarray(ref32, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    assert got.shape == ref64.shape, f"{what}: shape {got.shape} vs {ref64.shape}"
    row_err = np.max(np.abs(ref32 - ref64), axis=-1, keepdims=True)
    allowed = factor * row_err + TOL["f32"]["atol"] + TOL["f32"]["rtol"] * np.abs(ref64)
    err = np.abs(got - ref64)
    bad = ~(err <= allowed) & ~(np.isnan(got) & np.isnan(ref64))
    if bad.any():
        i = tuple(np.argwhere(bad)[0])
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} elements exceed {factor}x the reference's own "
                             f"float32 error; first at {i}: got {got[i]!r} want {ref64[i]!r} "
                             f"(allowed {allowed[i]:.3e}, err {err[i]:.3e})")
