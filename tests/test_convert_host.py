"""The converter recursions of csrc/convert_row.cuh, compiled for the HOST, against the oracle (CPU test).

The CUDA kernel (csrc/convert.cu) calls the same ``convert_row`` template on a shared-memory tile, so this pins
the recursion logic without a GPU; the GPU parity tests then cover the tiling and the launch.
"""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = {"lpc2par": 0, "par2lpc": 1, "gnorm": 2, "ignorm": 3, "norm0": 4}


@pytest.fixture(scope="module")
def lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = os.path.join(ROOT, "build", "tests")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libconvert_host.so")
    src = os.path.join(ROOT, "tests", "native", "convert_host.cu")
    # the device half of the shared headers is compiled too (and never run here): it needs the product's architecture
    subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler",
                    "-fPIC", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "diffsptk_b200", "csrc"), src, "-o", so], check=True)
    return C.CDLL(so)


CASES = [n for n in H.case_names() if H.load_case(n, "f64")[0] in OPS]


@pytest.mark.parametrize("fixed", [False, True], ids=["shared", "registers"])
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("name", CASES)
def test_host_build_of_the_kernel_recursion_matches_reference(lib, name, prec, fixed):
    from diffsptk_b200.utils import get_gamma

    op, params, ins, outs = H.load_case(name, prec)
    if fixed and op not in ("lpc2par", "par2lpc"):
        pytest.skip("register-resident variant exists for the O(M^2) recursions only")
    a = np.ascontiguousarray(ins[0]).copy()
    D = a.shape[-1]
    g = get_gamma(params.get("gamma", 0.0), params.get("c")) if op != "norm0" else 0.0
    fn = getattr(lib, f"convert_rows_{'fixed_' if fixed else ''}host_{prec}")
    fn(a.ctypes.data_as(C.c_void_p), C.c_long(a.size // D), C.c_int(D), C.c_int(OPS[op]), C.c_double(g))
    if prec == "f32" and op in ("lpc2par", "par2lpc"):
        # the step-down recursion divides by 1 - k^2: rows with |k| near 1 amplify float32 rounding, in the
        # reference as much as here (criterion of the other conditioning-dominated ops, see helpers)
        # ... measured against exact (float64 oracle) arithmetic on the SAME float32-rounded inputs: on these rows
        # the input rounding alone moves the result by more than any implementation difference
        exact = H.run_oracle(op, params, [ins[0].astype(np.float64)])
        H.assert_close_conditioned(a, outs[0], exact, what=f"{name}[f32]")
    else:
        H.assert_close(a, outs[0], prec, what=f"{name}[{prec}]", scale_atol=True)
