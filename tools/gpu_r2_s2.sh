#!/bin/bash
# round 2, GPU session 2: variant sweeps of the reworked kernels, the strict test-suite, ncu captures
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s2_build.txt 2>&1
: > $O/s2_sweep.jsonl
t() { env "$@" python tools/quick_time.py $WL 10 >> $O/s2_sweep.jsonl 2>> $O/s2_sweep.err; }
WL=stft;  t A=0; t DSB200_STFT_V=9; t DSB200_STFT_V=7; t DSB200_LIB_NAME=libdsb_legacy.so; t DSB200_STFT_V=0
WL=mfcc;  t A=0; t DSB200_MFCC_PLAN=0; t DSB200_MFCC_WARPS=12; t DSB200_LIB_NAME=libdsb_legacy.so DSB200_MFCC_PLAN=0
WL=mcep;  t DSB200_MCEP_V=12; t DSB200_MCEP_V=16; t DSB200_MCEP_V=8
WL=lpc;   t DSB200_LPC_W=12; t DSB200_LPC_W=8
WL=istft; t A=0; t DSB200_LIB_NAME=libdsb_legacy.so
WL=stft_grad; t A=0; t DSB200_LIB_NAME=libdsb_legacy.so
cat $O/s2_sweep.jsonl
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > $O/s2_pytest.txt 2>&1
tail -60 $O/s2_pytest.txt
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/s2_$1 python tools/prof_workload.py $2 4 > $O/s2_prof_$1.log 2>&1
  python tools/ncu_summary.py $O/s2_$1.ncu-rep $O/s2_ncu_$1 $4 >> $O/s2_prof_$1.log 2>&1
}
prof stft stft stft512_kernel 128000
prof mfcc mfcc stft512_kernel 512000
prof mcep mcep mcep_fast_kernel 1024000
prof lpc lpc lpc_wave_kernel 1024000
rm -f $O/s2_mcep.ncu-rep $O/s2_lpc.ncu-rep
cat $O/s2_ncu_*.txt
