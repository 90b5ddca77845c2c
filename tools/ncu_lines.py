"""Aggregate an ncu source page (sass+cuda) by CUDA source line: share of executed warp instructions, share of
stall samples and the top opcodes.  Usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    cur, agg, ops = None, collections.OrderedDict(), collections.defaultdict(collections.Counter)
    tot = tot_s = 0

    def num(v):
        try:
            return int(v)
        except ValueError:
            return 0
    for r in rows[h + 1:]:
        if len(r) <= iI:
            continue
        if r[0] != "":
            cur = (r[0], r[1].strip()[:60])
            agg.setdefault(cur, [0, 0])
            continue
        if cur is None:
            continue
        n, s = num(r[iI]), num(r[iS])
        agg[cur][0] += n
        agg[cur][1] += s
        tot += n
        tot_s += s
        tok = r[3].split()
        ops[cur][tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")] += n
    print(f"total warp instructions {tot}, samples {tot_s}")
    for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
        top = ", ".join(f"{o.split('.')[0]}:{c * 100 // max(n, 1)}" for o, c in ops[k].most_common(6))
        print(f"{k[0]:>4} {n / tot * 100:5.1f}% inst {s / max(tot_s, 1) * 100:5.1f}% smp | {k[1][:46]:46} | {top}")


if __name__ == "__main__":
    main()
