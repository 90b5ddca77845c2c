#!/bin/bash
# round 2, 8-GPU run of the bench line after the runs of quads (fused gather) and the SM margin (NCCL gather)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/n8_build.txt 2>&1
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 20 --warmup 5) > $O/n8_bench.json 2> $O/n8_bench.err
tail -c 2500 $O/n8_bench.json; tail -5 $O/n8_bench.err
