#!/bin/bash
# round 2, experiment session 4: lag-pair LPC kernel -- split Levinson chains, phase diagnostics
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/e4_build.txt 2>&1
(DSB200_LPC_V=52 DSB200_LPC_W2=16 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "lpc_wave or lpc_from") > $O/e4_pytest_v52.txt 2>&1
tail -n 3 $O/e4_pytest_v52.txt
(time timeout 600 python tools/sweep_knobs.py --steps 20 --out $O/e4_sweep.json \
  "lpc:LPC_V=20,52+LPC_W2=12,16" ) > $O/e4_sweep.txt 2> $O/e4_sweep.err
(DSB200_LIB_NAME=libdsb200_diag.so timeout 300 python tools/sweep_knobs.py --steps 20 --out $O/e4_diag.json \
  "lpc:LPC_V=84,212+LPC_W2=12,16") >> $O/e4_sweep.txt 2>> $O/e4_sweep.err
cat $O/e4_sweep.txt | cut -c1-260
tail -n 3 $O/e4_sweep.err
