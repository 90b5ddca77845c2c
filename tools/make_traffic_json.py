"""Write profiles/traffic.json (read by bench.py for roofline.traffic / roofline.fp32_pipe) from an ncu summary made by
tools/ncu_summary.py ON THE SAME SOURCE TREE: the entry carries the digest of the kernel sources, and bench.py drops it
as soon as those sources change.

    python tools/make_traffic_json.py stft profiles/r2_stft_final.json [out.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

SOURCES = {"stft": ["stft512.cu", "fft16.cuh", "bulk.cuh", "common.cuh"],
           "mfcc": ["stft512.cu", "fft16.cuh", "bulk.cuh", "common.cuh", "mfcc_plan.h"],
           "mcep": ["mcep_fast.cu", "common.cuh"], "lpc": ["fused_wave.cu", "bulk.cuh", "common.cuh"]}
PACKED = ("FADD2", "FFMA2", "FMUL2")
SCALAR = ("FADD", "FFMA", "FMUL")


def _us(m):
    scale = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}
    return None if not m else m["value"] * scale.get(m.get("unit", "us"), 1.0)


def main():
    wl, summary = sys.argv[1], json.load(open(sys.argv[2]))
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic.json")
    try:
        db = json.load(open(out))
    except Exception:
        db = {}
    per_unit = summary.get("warp_instructions_per_unit", {})
    units_are_quads = wl in ("stft", "mfcc")
    frames_per_unit = 4.0 if units_are_quads else 1.0
    slots = (sum(per_unit.get(k, 0.0) for k in PACKED) * 2 + sum(per_unit.get(k, 0.0) for k in SCALAR)) * 32
    db[wl] = {
        "dram_bytes_per_launch": summary.get("dram_traffic_bytes"),
        "capture": f"ncu --set full --clock-control none, one launch of {summary.get('kernel')} ({os.path.basename(sys.argv[2])})",
        "kernel_us_under_ncu": _us(summary["metrics"].get("gpu__time_duration.sum", {})),
        "fma_lane_slots_per_frame": slots / frames_per_unit if slots else None,
        "sources": SOURCES[wl],
        "source_digest": bench.kernel_source_digest(SOURCES[wl]),
    }
    with open(out, "w") as f:
        json.dump(db, f, indent=1, sort_keys=True)
    print(json.dumps(db[wl]))


if __name__ == "__main__":
    main()
