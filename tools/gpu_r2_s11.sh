#!/bin/bash
# round 2, GPU session 11 (2 GPUs): gathers after the SM margin (NCCL) and the runs of quads (fused)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/s11_build.txt 2>&1
(time timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider) > $O/s11_pytest_multi.txt 2>&1
tail -5 $O/s11_pytest_multi.txt
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 2 --steps 20 --warmup 5) > $O/s11_bench_n2.json 2> $O/s11_bench_n2.err
python -c "
import json
j=json.loads(open('$O/s11_bench_n2.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'])
for k,v in j['extra_workloads'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='what'})
"; tail -3 $O/s11_bench_n2.err
