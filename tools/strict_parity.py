"""Every float32 golden case at the reference's UNSCALED tolerance (rtol 1e-4, atol 1e-6; tests/utils.py:66-72 of the
reference), on the GPU box.  For each case three comparisons with that same criterion:

    ours vs reference-f32   (what the parity tests assert)
    ours vs reference-f64   (distance to the truth)
    reference-f32 vs reference-f64   (the reference's own float32 noise: a case that fails HERE cannot be held to
                                      the unscaled tolerance by any float32 implementation)

Writes a table to stdout and the list of cases that need a relaxed criterion to
tests/golden/strict_exceptions.json (only with --write).

    python tools/strict_parity.py [--write] > profiles/r2_accuracy.txt
"""
import collections
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402

import diffsptk_b200.functional as F  # noqa: E402

RTOL, ATOL = 1e-4, 1e-6


def as_real(a):
    a = np.asarray(a)
    if np.iscomplexobj(a):
        a = np.stack([a.real, a.imag], -1)
    return a.astype(np.float64)


def n_bad(a, b):
    a, b = as_real(a), as_real(b)
    return int((~np.isclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True)).sum()), a.size


def main():
    rows, exc = [], {}
    per_op = collections.defaultdict(lambda: [0, 0, 0, 0])
    for name in H.case_names():
        op, params, ins32, out32 = H.load_case(name, "f32")
        _, _, _, out64 = H.load_case(name, "f64")
        with torch.no_grad():
            ins = [None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in ins32]
            if op == "mgcep":
                import diffsptk_b200 as B
                got = B.MelGeneralizedCepstralAnalysis(**params).cuda()(*ins)
            else:
                got = getattr(F, op)(*ins, **params)
        got = got if isinstance(got, tuple) else (got,)
        b32 = b64 = r = tot = 0
        for g, r32, r64 in zip(got, out32, out64):
            g = g.cpu().numpy()
            x, n = n_bad(g, r32); b32 += x; tot += n
            b64 += n_bad(g, r64)[0]
            r += n_bad(r32, r64)[0]
        s = per_op[op]
        s[0] += 1; s[1] += b32 > 0; s[2] += b64 > 0; s[3] += r > 0
        if b32 or b64 or r:
            rows.append((name, op, tot, b32, b64, r))
        if b32:
            exc[name] = {"op": op, "elements": tot, "ours_vs_ref32": b32, "ours_vs_ref64": b64, "ref32_vs_ref64": r}
    print(f"# unscaled rtol {RTOL:g} / atol {ATOL:g}; counts are elements outside the tolerance")
    print(f"{'op':10s} {'cases':>5s} {'ours!=ref32':>12s} {'ours!=ref64':>12s} {'ref32!=ref64':>13s}   (cases with any outlier)")
    for op in sorted(per_op):
        n, a, b, c = per_op[op]
        print(f"{op:10s} {n:5d} {a:12d} {b:12d} {c:13d}")
    print()
    print(f"{'case':40s} {'op':9s} {'elements':>9s} {'ours!=ref32':>12s} {'ours!=ref64':>12s} {'ref32!=ref64':>13s}")
    for name, op, tot, b32, b64, r in rows:
        print(f"{name:40s} {op:9s} {tot:9d} {b32:12d} {b64:12d} {r:13d}")
    if "--write" in sys.argv:
        with open(os.path.join(ROOT, "tests", "golden", "strict_exceptions.json"), "w") as f:
            json.dump(exc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
