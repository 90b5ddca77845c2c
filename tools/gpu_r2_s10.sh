#!/bin/bash
# round 2, GPU session 10: stftn with the interior staging path; full suite
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s10_build.txt 2>&1
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > $O/s10_pytest.txt 2>&1
tail -8 $O/s10_pytest.txt
: > $O/s10_sweep.jsonl
t() { env "$@" python tools/quick_time.py $WL 10 >> $O/s10_sweep.jsonl 2>> $O/s10_sweep.err; }
WL=stft;      t A=0
WL=stft1024;  t A=0
WL=stft2048;  t A=0
cat $O/s10_sweep.jsonl
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/s10_$1 python tools/prof_workload.py $2 4 > $O/s10_prof_$1.log 2>&1
  python tools/ncu_summary.py $O/s10_$1.ncu-rep $O/s10_ncu_$1 $4 >> $O/s10_prof_$1.log 2>&1
  python tools/ncu_lines.py $O/s10_$1.ncu-rep 30 > $O/s10_lines_$1.txt 2>&1
  rm -f $O/s10_$1.ncu-rep
}
prof stft1024 stft1024 stftn_kernel 128000
cat $O/s10_ncu_stft1024.txt; head -30 $O/s10_lines_stft1024.txt
