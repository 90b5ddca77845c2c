"""CUDA-event timing of one bench workload, nothing else (tuning knobs are read from the environment once per
process, so every variant is one invocation):  python tools/quick_time.py WORKLOAD [steps]  ->  one JSON line."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    wl = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda", 0)
    _, B, T, rd, wr = bench.WORKLOADS[wl]
    xs, step = bench.make_step(wl, B, T, dev)
    _, per = bench.timed_steps(step, steps, 3, False)
    ms = statistics.mean(per)
    frames = B * bench.n_frames(T, bench._HOPS.get(wl, bench.FP))
    knobs = {k: v for k, v in os.environ.items() if k.startswith("DSB200_")}
    print(json.dumps({"workload": wl, "ms": ms, "min_ms": min(per), "frames_per_s": frames / (ms / 1e3),
                      "hbm_frac": frames * (rd + wr) / (ms / 1e3) / 1e9 / bench.hbm_peak()[0], "knobs": knobs}))


if __name__ == "__main__":
    main()
