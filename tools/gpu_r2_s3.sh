#!/bin/bash
# round 2, GPU session 3: ncu captures of the reworked kernels, strict-parity evidence incl. BASELINE shapes, tests
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s3_build.txt 2>&1
prof() {  # name workload kernel-regex units
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -f -o $O/s3_$1 python tools/prof_workload.py $2 4 > $O/s3_prof_$1.log 2>&1
  python tools/ncu_summary.py $O/s3_$1.ncu-rep $O/s3_ncu_$1 $4 >> $O/s3_prof_$1.log 2>&1
  python tools/ncu_lines.py $O/s3_$1.ncu-rep 45 > $O/s3_lines_$1.txt 2>&1
}
prof mcep mcep mcep_fast_kernel 1024000
prof mfcc mfcc stft512_kernel 512000
prof lpc lpc lpc_wave_kernel 1024000
prof stft stft stft512_kernel 128000
cat $O/s3_ncu_*.txt
python tools/strict_parity.py > $O/s3_strict.txt 2> $O/s3_strict.err
tail -25 $O/s3_strict.txt
(time python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > $O/s3_pytest.txt 2>&1
tail -15 $O/s3_pytest.txt
ls -la $O/*.ncu-rep
