#!/bin/bash
# round 2, GPU session 7: widened stft512 envelope, mcep default + rows4 (bank-conflict fix), tests, final ncu + traffic.json
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s7_build.txt 2>&1
: > $O/s7_sweep.jsonl
t() { env "$@" python tools/quick_time.py $WL 10 >> $O/s7_sweep.jsonl 2>> $O/s7_sweep.err; }
WL=stft;  t A=0
WL=mcep;  t DSB200_MCEP_V=12; t DSB200_MCEP_V=122; t DSB200_MCEP_V=16
WL=stft;  t A=1
cat $O/s7_sweep.jsonl
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > $O/s7_pytest.txt 2>&1
tail -15 $O/s7_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/s7_smoke.txt 2>&1; tail -3 $O/s7_smoke.txt
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/s7_$1 python tools/prof_workload.py $2 4 > $O/s7_prof_$1.log 2>&1
  python tools/ncu_summary.py $O/s7_$1.ncu-rep $O/s7_ncu_$1 $4 >> $O/s7_prof_$1.log 2>&1
  python tools/ncu_lines.py $O/s7_$1.ncu-rep 40 > $O/s7_lines_$1.txt 2>&1
  python tools/make_traffic_json.py $2 $O/s7_ncu_$1.json $O/s7_traffic.json >> $O/s7_prof_$1.log 2>&1
  rm -f $O/s7_$1.ncu-rep
}
prof stft stft stft512_kernel 128000
prof mfcc mfcc stft512_kernel 512000
prof mcep mcep mcep_fast_kernel 1024000
prof lpc lpc lpc_wave_kernel 1024000
cat $O/s7_traffic.json
# launch list of the default bench (shares of the step, cold-cache and serialised under ncu)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s7_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > $O/s7_launches_bench.log 2>&1
python - <<'PY' > $O/s7_launches_summary.txt 2>&1
import csv, collections
rows = list(csv.reader(open('gpurun_out/s7_launches.csv')))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[h]; iK = hdr.index('Kernel Name'); iV = hdr.index('Metric Value'); iU = hdr.index('Metric Unit')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= iV: continue
    v = float(r[iV].replace(',', '')); u = r[iU]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(u, 1.0)
    k = r[iK][:90]; tot[k] += v; cnt[k] += 1
s = sum(tot.values())
print(f"total {s:.1f} us over {sum(cnt.values())} launches")
for k, v in tot.most_common(12): print(f"{v / s * 100:6.2f} %  {cnt[k]:4d} x  {v / cnt[k]:9.1f} us  {k}")
PY
cat $O/s7_launches_summary.txt
(time python bench.py --steps 20 --warmup 5) > $O/s7_bench.json 2> $O/s7_bench.err
head -c 1200 $O/s7_bench.json
