"""One process, many kernel variants: sets the library's tuning knobs through ``dsb200_set_knob`` (looked up at every
launch), checks each variant's output against the default build's on the same input, and times it with CUDA events.

    python tools/sweep_knobs.py [--steps 20] [--out gpurun_out/sweep.json] SPEC [SPEC ...]

SPEC = ``workload:KNOB=v1,v2,...[+KNOB2=w1,w2,...]`` -- the cross product of the listed values, e.g.
``lpc:LPC_STAGGER=0,8000,16000+LPC_V=0,7``.  Knob names are README.md's without the ``DSB200_`` prefix."""
import argparse
import itertools
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_spec(spec):
    wl, rest = spec.split(":", 1)
    axes = []
    for part in rest.split("+"):
        name, vals = part.split("=")
        axes.append((name, [int(v) for v in vals.split(",")]))
    names = [a[0] for a in axes]
    return wl, [dict(zip(names, combo)) for combo in itertools.product(*[a[1] for a in axes])]


def main():
    import torch

    import bench
    from diffsptk_b200 import _native
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--repeat", type=int, default=2, help="timed passes per setting (the best mean is reported)")
    ap.add_argument("--out", default=None)
    ap.add_argument("specs", nargs="+")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    for spec in a.specs:
        wl, settings = parse_spec(spec)
        _, B, T, rd, wr = bench.WORKLOADS[wl]
        xs, step = bench.make_step(wl, B, T, dev)
        frames = B * bench.n_frames(T, bench._HOPS.get(wl, bench.FP))
        _native.clear_knobs()
        with torch.no_grad():
            y0 = step(0).clone()
        kernel0 = _native.last_kernel()
        for knobs in [{}] + settings:
            _native.clear_knobs()
            for k, v in knobs.items():
                _native.set_knob(k, v)
            try:
                with torch.no_grad():
                    y = step(0)
                torch.cuda.synchronize()
                diff = float((y.double() - y0.double()).abs().max())
                bad = int((~torch.isfinite(y)).sum())
                means, mins = [], []
                for _ in range(a.repeat):
                    _, per = bench.timed_steps(step, a.steps, 3, False)
                    means.append(statistics.mean(per))
                    mins.append(min(per))
                row = {"workload": wl, "knobs": knobs, "ms": min(means), "min_ms": min(mins),
                       "frames_per_s": frames / (min(means) / 1e3), "max_abs_diff_vs_default": diff,
                       "non_finite": bad, "kernel": _native.last_kernel()}
            except Exception as e:  # noqa: BLE001 -- a variant that does not launch is a result, not a crash
                row = {"workload": wl, "knobs": knobs, "error": str(e)[:200]}
            rows.append(row)
            print(json.dumps(row), flush=True)
        _native.clear_knobs()
        del xs, step, y0
        torch.cuda.empty_cache()
        assert kernel0
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
