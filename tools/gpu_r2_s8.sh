#!/bin/bash
# round 2, GPU session 8 (8 GPUs): the scaling line incl. the gathers at N = 8, and N = 4
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s8_build.txt 2>&1
for N in 8 4; do
  (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N)) bench.py --gpus $N --steps 20 --warmup 5) > $O/s8_bench_n$N.json 2> $O/s8_bench_n$N.err
  tail -c 2500 $O/s8_bench_n$N.json; tail -5 $O/s8_bench_n$N.err
done
(time timeout 300 python bench.py --impl reference --gpus 8 --steps 20 --warmup 5) > $O/s8_ref.json 2> $O/s8_ref.err
cat $O/s8_ref.json
