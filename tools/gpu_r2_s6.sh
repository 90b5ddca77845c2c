#!/bin/bash
# round 2, GPU session 6: mcep solve variants (four rows per lane vs lane = row; phased vs exit-test loops)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s6_build.txt 2>&1
: > $O/s6_sweep.jsonl
t() { env "$@" python tools/quick_time.py $WL 10 >> $O/s6_sweep.jsonl 2>> $O/s6_sweep.err; }
WL=mcep;  t DSB200_MCEP_V=12; t DSB200_MCEP_V=16; t DSB200_MCEP_V=121; t DSB200_MCEP_V=161; t DSB200_LIB_NAME=libdsb_break.so DSB200_MCEP_V=121
cat $O/s6_sweep.jsonl
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "mcep or mgcep or smoke") > $O/s6_pytest.txt 2>&1
tail -8 $O/s6_pytest.txt
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/s6_$1 python tools/prof_workload.py $2 4 > $O/s6_prof_$1.log 2>&1
  python tools/ncu_summary.py $O/s6_$1.ncu-rep $O/s6_ncu_$1 $4 >> $O/s6_prof_$1.log 2>&1
  python tools/ncu_lines.py $O/s6_$1.ncu-rep 40 > $O/s6_lines_$1.txt 2>&1
  rm -f $O/s6_$1.ncu-rep
}
prof mcep mcep mcep_fast_kernel 1024000
cat $O/s6_ncu_mcep.txt
