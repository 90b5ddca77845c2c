"""Tiny driver for ncu: a few steps of one bench workload (python tools/prof_workload.py WORKLOAD [steps])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402

wl = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
_, B, T, _, _ = bench.WORKLOADS[wl]
xs, step = bench.make_step(wl, B, T, torch.device("cuda", 0))
with torch.no_grad():
    for i in range(n):
        y = step(i)
torch.cuda.synchronize()
print(float(y.float().sum()))
