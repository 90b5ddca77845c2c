#!/bin/bash
# round 2, GPU session 9: stftn (fft_length 1024 / 2048), MFCC runs of quads, full suite
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s9_build.txt 2>&1
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > $O/s9_pytest.txt 2>&1
tail -25 $O/s9_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/s9_smoke.txt 2>&1; tail -3 $O/s9_smoke.txt
: > $O/s9_sweep.jsonl
t() { env "$@" python tools/quick_time.py $WL 10 >> $O/s9_sweep.jsonl 2>> $O/s9_sweep.err; }
WL=stft;      t A=0
WL=stft1024;  t A=0; t DSB200_STFT_GENERIC=1
WL=stft2048;  t A=0; t DSB200_STFT_GENERIC=1
WL=mfcc;      t A=0; t DSB200_MFCC_RUN=1
WL=mcep;      t A=0
cat $O/s9_sweep.jsonl
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/s9_$1 python tools/prof_workload.py $2 4 > $O/s9_prof_$1.log 2>&1
  python tools/ncu_summary.py $O/s9_$1.ncu-rep $O/s9_ncu_$1 $4 >> $O/s9_prof_$1.log 2>&1
  python tools/ncu_lines.py $O/s9_$1.ncu-rep 30 > $O/s9_lines_$1.txt 2>&1
  rm -f $O/s9_$1.ncu-rep
}
prof stft1024 stft1024 stftn_kernel 128000
prof mfcc mfcc stft512_kernel 512000
python tools/make_traffic_json.py mfcc $O/s9_ncu_mfcc.json $O/s9_traffic.json >> $O/s9_prof_mfcc.log 2>&1
cat $O/s9_ncu_stft1024.txt
