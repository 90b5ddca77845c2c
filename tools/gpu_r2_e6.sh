#!/bin/bash
# round 2, session 6: compute-sanitizer over the extended smoke (lag-pair LPC kernel, shared-memory FFT kernel,
# zmean / relative-floor builds included); the MFCC tests alone (launch-count assertions without an earlier test)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/e6_build.txt 2>&1
python tools/sanitize_smoke.py > $O/e6_smoke_plain.txt 2>&1; tail -n 2 $O/e6_smoke_plain.txt
for tool in memcheck racecheck synccheck; do
  (time timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=dsb200 --print-limit 20 python tools/sanitize_smoke.py) > $O/e6_san_$tool.txt 2>&1
  tail -n 6 $O/e6_san_$tool.txt
done
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "mfcc") > $O/e6_pytest_mfcc.txt 2>&1
tail -n 3 $O/e6_pytest_mfcc.txt
