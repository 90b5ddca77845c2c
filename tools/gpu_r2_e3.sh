#!/bin/bash
# round 2, experiment session 3: lag-pair form of the LPC kernel (scalar-broadcast FFMA2), 12 and 16 warps
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/e3_build.txt 2>&1
(DSB200_LPC_V=20 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "lpc_wave or lpc_from") > $O/e3_pytest_v20.txt 2>&1
tail -n 3 $O/e3_pytest_v20.txt
(DSB200_LPC_V=20 DSB200_LPC_W2=16 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "lpc_wave or lpc_from") > $O/e3_pytest_v20w16.txt 2>&1
tail -n 3 $O/e3_pytest_v20w16.txt
(time timeout 600 python tools/sweep_knobs.py --steps 20 --out $O/e3_sweep.json \
  "lpc:LPC_V=7,16,20+LPC_W2=12,16" ) > $O/e3_sweep.txt 2> $O/e3_sweep.err
cat $O/e3_sweep.txt | cut -c1-260
tail -n 3 $O/e3_sweep.err
