#!/bin/bash
# round 2, GPU session 4 (2 GPUs): gather paths -- NCCL in place, fused into the kernel -- correctness and timing
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi topo -m > $O/s4_topo.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/s4_build.txt 2>&1
python tools/quick_time.py mfcc 10 > $O/s4_mfcc_1gpu.json 2> $O/s4_mfcc_1gpu.err; cat $O/s4_mfcc_1gpu.json
(time timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider) > $O/s4_pytest_multi.txt 2>&1
tail -30 $O/s4_pytest_multi.txt
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5) > $O/s4_bench_n2.json 2> $O/s4_bench_n2.err
cat $O/s4_bench_n2.json; tail -20 $O/s4_bench_n2.err
DSB200_GATHER_MODE=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 > $O/s4_bench_n2_p2p.json 2> $O/s4_bench_n2_p2p.err
cat $O/s4_bench_n2_p2p.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(json.dumps(j['extra_workloads'], indent=1))"
