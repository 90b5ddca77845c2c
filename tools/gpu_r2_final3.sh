#!/bin/bash
# round 2, closing session: full suite + smoke on the shipped build, re-capture of the two workloads whose kernel source
# changed since the last captures (stft512.cu: staging-buffer padding), bench lines
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/h_build.txt 2>&1
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > $O/h_pytest.txt 2>&1
tail -n 6 $O/h_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/h_smoke.txt 2>&1; tail -n 2 $O/h_smoke.txt
: > $O/h_sweep.jsonl
for WL in stft mfcc mcep lpc stft1024 stft2048 istft stft_grad; do python tools/quick_time.py $WL 10 >> $O/h_sweep.jsonl 2>> $O/h_sweep.err; done
cat $O/h_sweep.jsonl | cut -c1-160
cp profiles/traffic.json $O/h_traffic.json
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/h_$1 python tools/prof_workload.py $2 4 > $O/h_prof_$1.log 2>&1
  python tools/ncu_summary.py $O/h_$1.ncu-rep $O/h_ncu_$1 $4 >> $O/h_prof_$1.log 2>&1
  python tools/ncu_lines.py $O/h_$1.ncu-rep 40 > $O/h_lines_$1.txt 2>&1
  python tools/make_traffic_json.py $2 $O/h_ncu_$1.json $O/h_traffic.json >> $O/h_prof_$1.log 2>&1
  [ "$1" = "stft" ] || rm -f $O/h_$1.ncu-rep
}
prof stft stft stft512_kernel 128000
prof mfcc mfcc stft512_kernel 512000
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/h_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > $O/h_launches_bench.log 2>&1
(time python bench.py --steps 20 --warmup 5) > $O/h_bench.json 2> $O/h_bench.err
head -c 1200 $O/h_bench.json
