#!/bin/bash
# round 2, GPU session 5: rolled mcep solve, 12-warp PAIR2 MFCC, lpc with out-of-line staging; tests; bench
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s5_build.txt 2>&1
: > $O/s5_sweep.jsonl
t() { env "$@" python tools/quick_time.py $WL 10 >> $O/s5_sweep.jsonl 2>> $O/s5_sweep.err; }
WL=mcep;  t DSB200_MCEP_V=12; t DSB200_MCEP_V=16; t DSB200_MCEP_V=120
WL=mfcc;  t A=0; t DSB200_MFCC_WARPS=16
WL=lpc;   t A=0; t DSB200_LPC_W=8
cat $O/s5_sweep.jsonl
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > $O/s5_pytest.txt 2>&1
tail -15 $O/s5_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/s5_smoke.txt 2>&1; tail -3 $O/s5_smoke.txt
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/s5_$1 python tools/prof_workload.py $2 4 > $O/s5_prof_$1.log 2>&1
  python tools/ncu_summary.py $O/s5_$1.ncu-rep $O/s5_ncu_$1 $4 >> $O/s5_prof_$1.log 2>&1
  python tools/ncu_lines.py $O/s5_$1.ncu-rep 40 > $O/s5_lines_$1.txt 2>&1
  rm -f $O/s5_$1.ncu-rep
}
prof mcep mcep mcep_fast_kernel 1024000
prof mfcc mfcc stft512_kernel 512000
cat $O/s5_ncu_mcep.txt $O/s5_ncu_mfcc.txt
(time python bench.py --steps 20 --warmup 5) > $O/s5_bench.json 2> $O/s5_bench.err
head -c 1500 $O/s5_bench.json
