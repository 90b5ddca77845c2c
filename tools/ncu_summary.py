"""Summarise an .ncu-rep (one kernel, `ncu --set full --import-source on`) into a small JSON + text file
for `profiles/`:  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name [units_per_launch]

Reads the raw page (device-level metrics) and the source page (per-SASS-instruction counts and stall
samples) through `ncu -i ... --csv`, which works on the CPU-only build box.
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
]
STALLS = ["stall_barrier", "stall_lg", "stall_long_sb", "stall_math", "stall_mio", "stall_short_sb", "stall_wait",
          "stall_not_selected", "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst",
          "stall_sleep", "stall_membar", "stall_drain"]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = page(rep, "raw")
    hdr, unit, val = raw[0], raw[1], raw[2]
    summary = {"kernel": val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?", "metrics": {}}
    for h, u, v in zip(hdr, unit, val):
        if h in RAW:
            summary["metrics"][h] = {"value": num(v), "unit": u}
    src = page(rep, "source", ["--print-source", "sass"])
    shdr, data = src[1], src[2:]
    c = {h: i for i, h in enumerate(shdr)}
    stall = collections.Counter()
    ops = collections.Counter()
    total = 0
    for r in data:
        for s in STALLS:
            if s in c:
                stall[s] += int(r[c[s]])
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[c["Source"]].strip())
        n = int(r[c["Instructions Executed"]])
        ops[m.group(2) if m else "?"] += n
        total += n
    tot_s = sum(stall.values()) or 1
    summary["stall_pct"] = {k: round(100.0 * v / tot_s, 1) for k, v in stall.most_common() if v}
    summary["warp_instructions"] = total
    top = ops.most_common(24)
    summary["warp_instructions_by_opcode"] = {k: v for k, v in top}
    if units:
        summary["units_per_launch"] = units
        summary["warp_instructions_per_unit"] = {k: round(v / units, 2) for k, v in top}
        summary["warp_instructions_per_unit_total"] = round(total / units, 1)
    m = summary["metrics"]
    rd, wr = m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum")
    if rd and wr:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        summary["dram_traffic_bytes"] = rd["value"] * scale.get(rd["unit"], 1) + wr["value"] * scale.get(wr["unit"], 1)
    with open(out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    with open(out + ".txt", "w") as f:
        f.write(f"kernel: {summary['kernel']}\n")
        for k, v in m.items():
            f.write(f"{k:72s} {v['value']} {v['unit']}\n")
        f.write("stall %: " + json.dumps(summary["stall_pct"]) + "\n")
        if units:
            f.write(f"warp instructions per unit ({units:g} units): {summary['warp_instructions_per_unit_total']}\n")
            f.write(json.dumps(summary["warp_instructions_per_unit"]) + "\n")
    print(out + ".json")


if __name__ == "__main__":
    main()
