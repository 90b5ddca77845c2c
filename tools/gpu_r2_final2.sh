#!/bin/bash
# round 2, final single-GPU session (second: after the LPC lag-pair kernel, knobs per launch, mcep rows-of-four default): full suite, smoke, final ncu captures + traffic.json, launch list, bench line
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/g_build.txt 2>&1
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > $O/g_pytest.txt 2>&1
tail -12 $O/g_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/g_smoke.txt 2>&1; tail -3 $O/g_smoke.txt
: > $O/g_sweep.jsonl
t() { env "$@" python tools/quick_time.py $WL 10 >> $O/g_sweep.jsonl 2>> $O/g_sweep.err; }
for WL in stft mfcc mcep lpc stft1024 stft2048 istft stft_grad; do t A=0; done
cat $O/g_sweep.jsonl
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/g_$1 python tools/prof_workload.py $2 4 > $O/g_prog_$1.log 2>&1
  python tools/ncu_summary.py $O/g_$1.ncu-rep $O/g_ncu_$1 $4 >> $O/g_prog_$1.log 2>&1
  python tools/ncu_lines.py $O/g_$1.ncu-rep 40 > $O/g_lines_$1.txt 2>&1
  python tools/make_traffic_json.py $2 $O/g_ncu_$1.json $O/g_traffic.json >> $O/g_prog_$1.log 2>&1
  [ "$1" = "stft" ] || rm -f $O/g_$1.ncu-rep
}
prof stft stft stft512_kernel 128000
prof mfcc mfcc stft512_kernel 512000
prof mcep mcep mcep_fast_kernel 1024000
prof lpc lpc lpc_wave 1024000
cat $O/g_traffic.json | head -60
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/g_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > $O/g_launches_bench.log 2>&1
(time python bench.py --steps 20 --warmup 5) > $O/g_bench.json 2> $O/g_bench.err
head -c 1500 $O/g_bench.json
(time python bench.py --impl reference --steps 20 --warmup 5) > $O/g_ref.json 2> $O/g_ref.err
head -c 600 $O/g_ref.json
