#!/bin/bash
# round 2, experiment session 2: FFMA2 issue-rate micro-benchmark (operand reuse), rolled Levinson in the LPC kernel
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/e2_build.txt 2>&1
tools/bin/bench_ffma2 > $O/e2_ffma2.jsonl 2>&1
cat $O/e2_ffma2.jsonl
(time timeout 600 python tools/sweep_knobs.py --steps 20 --out $O/e2_sweep.json \
  "lpc:LPC_V=0,3,7,8,11" ) > $O/e2_sweep.txt 2> $O/e2_sweep.err
(DSB200_LIB_NAME=libdsb200_diag.so timeout 300 python tools/sweep_knobs.py --steps 20 --out $O/e2_diag.json \
  "lpc:LPC_V=64,192,195") >> $O/e2_sweep.txt 2>> $O/e2_sweep.err
cat $O/e2_sweep.txt | cut -c1-260
tail -3 $O/e2_sweep.err
(time timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "lpc") > $O/e2_pytest.txt 2>&1
tail -3 $O/e2_pytest.txt
(DSB200_LPC_V=11 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "lpc") > $O/e2_pytest_v11.txt 2>&1
tail -3 $O/e2_pytest_v11.txt
