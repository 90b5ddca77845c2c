#!/bin/bash
# round 2, experiment session 5: MFCC with the window / split twiddles in registers; mcep 16-warp rows-of-four
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export DSB200_LIB_NAME=libdsb200_rt.so
(time timeout 600 python tools/sweep_knobs.py --steps 20 --out $O/e5_sweep.json \
  "mfcc:MFCC_RT=0,1" "mfcc:MFCC_RT=0,1" "mcep:MCEP_V=122,162,12" ) > $O/e5_sweep.txt 2> $O/e5_sweep.err
cat $O/e5_sweep.txt | cut -c1-260
tail -n 3 $O/e5_sweep.err
(DSB200_MFCC_RT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "mfcc") > $O/e5_pytest_rt.txt 2>&1
tail -n 3 $O/e5_pytest_rt.txt
