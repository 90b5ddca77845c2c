#!/bin/bash
# round 2, experiment session 1: start-up stagger of the warps sharing a scheduler (all four persistent kernels),
# LPC variants (halo through the window's zero tail, packed window product, order-24 Levinson without guards)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/e1_build.txt 2>&1
(time timeout 600 python tools/sweep_knobs.py --steps 20 --out $O/e1_sweep.json \
  "lpc:LPC_V=0,1,2,4,7" \
  "lpc:LPC_STAGGER=4000,8000,12000,18000,25000,35000+LPC_V=0,7" \
  "lpc:LPC_W=8+LPC_STAGGER=0,12000,25000,40000" \
  "stft:STFT_STAGGER=400,800,1600,3000" \
  "mfcc:MFCC_STAGGER=1000,3000,6000,12000,24000" \
  "mcep:MCEP_STAGGER=10000,20000,44000,80000" \
  "mfcc:MFCC_WARPS=16+MFCC_STAGGER=0,2000,5000" \
  "mcep:MCEP_V=122+MCEP_STAGGER=0,20000,44000" ) > $O/e1_sweep.txt 2> $O/e1_sweep.err
# timing diagnostics of the LPC kernel (wrong results by construction): no Levinson / no reduction / neither
(DSB200_LIB_NAME=libdsb200_diag.so timeout 300 python tools/sweep_knobs.py --steps 20 --out $O/e1_diag.json \
  "lpc:LPC_V=8,16,24+LPC_STAGGER=0,12000") >> $O/e1_sweep.txt 2>> $O/e1_sweep.err
cat $O/e1_sweep.txt | cut -c1-260
tail -3 $O/e1_sweep.err
(time timeout 300 python -m pytest tests/test_gpu_autograd.py -q -x -p no:cacheprovider -k "learnable_dft_basis") > $O/e1_pytest.txt 2>&1
tail -3 $O/e1_pytest.txt
