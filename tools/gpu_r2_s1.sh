#!/bin/bash
# round 2, GPU session 1: baseline of the untouched kernels -- tests, strict parity table, sanitizer, new bench.py
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi -L > $O/s1_gpus.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/s1_build.txt 2>&1
(time python -m pytest tests -m gpu -x -q) > $O/s1_pytest.txt 2>&1
python tools/strict_parity.py --write > $O/s1_strict.txt 2> $O/s1_strict.err
cp tests/golden/strict_exceptions.json $O/s1_strict_exceptions.json
(time python bench.py --steps 20 --warmup 5) > $O/s1_bench.json 2> $O/s1_bench.err
for tool in memcheck racecheck synccheck initcheck; do
  (time timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=dsb200 --print-limit 20 python tools/sanitize_smoke.py) > $O/s1_san_$tool.txt 2>&1
  tail -5 $O/s1_san_$tool.txt
done
tail -3 $O/s1_pytest.txt; tail -30 $O/s1_strict.txt; cat $O/s1_bench.json | head -c 3000
