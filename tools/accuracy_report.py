"""Accuracy of the fp32 kernels relative to the reference's own fp32 error (run on the GPU box).

For every golden case the committed fixtures hold the reference's float32 and float64 outputs for the same
inputs.  This script runs the float32 kernels on the float32 inputs and reports, per op,

    err_ref  = max |ref_f32 - ref_f64|          (the reference's own rounding noise)
    err_ours = max |ours_f32 - ref_f64|

both normalised by max |ref_f64| of the case, plus the worst ratio err_ours / err_ref.  A ratio near or below 1
means the kernels are as close to the float64 truth as the reference itself is.

    python tools/accuracy_report.py > profiles/r1_accuracy.txt
"""
import collections
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402

import diffsptk_b200.functional as F  # noqa: E402


def to_dev(a):
    if a is None:
        return None
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.cuda()


def main():
    stats = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0, ""])
    for name in H.case_names():
        op, params, ins32, out32 = H.load_case(name, "f32")
        _, _, _, out64 = H.load_case(name, "f64")
        with torch.no_grad():
            if op == "mgcep":   # an nn.Module without functional form, in the reference too
                import diffsptk_b200 as B
                got = B.MelGeneralizedCepstralAnalysis(**params).cuda()(*[to_dev(a) for a in ins32])
            else:
                got = getattr(F, op)(*[to_dev(a) for a in ins32], **params)
        got = got if isinstance(got, tuple) else (got,)
        for g, r32, r64 in zip(got, out32, out64):
            g = g.cpu().numpy()
            r64c = r64.astype(np.complex128 if np.iscomplexobj(r64) else np.float64)
            fin = np.isfinite(r64c) & np.isfinite(r32) & np.isfinite(g)
            if not fin.any():
                continue
            scale = max(float(np.max(np.abs(r64c[fin]))), 1e-30)
            e_ref = float(np.max(np.abs(r32[fin].astype(r64c.dtype) - r64c[fin]))) / scale
            e_our = float(np.max(np.abs(g[fin].astype(r64c.dtype) - r64c[fin]))) / scale
            s = stats[op]
            s[0], s[1] = max(s[0], e_ref), max(s[1], e_our)
            ratio = e_our / max(e_ref, 1e-9)
            if ratio > s[2]:
                s[2], s[4] = ratio, name
            s[3] += 1
    print(f"{'op':10s} {'cases':>5s} {'max err_ref':>12s} {'max err_ours':>13s} {'worst ours/ref':>15s}  worst case")
    for op in sorted(stats):
        e_ref, e_our, ratio, n, worst = stats[op]
        print(f"{op:10s} {n:5d} {e_ref:12.3e} {e_our:13.3e} {ratio:15.2f}  {worst}")


if __name__ == "__main__":
    main()
