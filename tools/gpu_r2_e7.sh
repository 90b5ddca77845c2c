#!/bin/bash
# round 2, session 7: stftn with register-prefetched staging (tests + timing); racecheck / memcheck over shapes that keep
# every persistent warp in its main loop for several iterations, without and with the staging-buffer padding
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/e7_build.txt 2>&1
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "large_fft") > $O/e7_pytest_stftn.txt 2>&1
tail -n 3 $O/e7_pytest_stftn.txt
: > $O/e7_time.jsonl
for WL in stft1024 stft2048 stft mfcc; do python tools/quick_time.py $WL 10 >> $O/e7_time.jsonl 2>> $O/e7_time.err; done
DSB200_STFT_INPAD=16 python tools/quick_time.py stft 10 >> $O/e7_time.jsonl 2>> $O/e7_time.err
DSB200_STFT_INPAD=16 python tools/quick_time.py mfcc 10 >> $O/e7_time.jsonl 2>> $O/e7_time.err
cat $O/e7_time.jsonl | cut -c1-200
python tools/sanitize_loop.py > $O/e7_loop_plain.txt 2>&1; tail -n 2 $O/e7_loop_plain.txt
(time timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=dsb200 --print-limit 30 python tools/sanitize_loop.py) > $O/e7_race_pad0.txt 2>&1
grep -E "Warning|Error|SUMMARY|ok" $O/e7_race_pad0.txt | cut -c1-260 | head -20
(time DSB200_STFT_INPAD=16 timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=dsb200 --print-limit 30 python tools/sanitize_loop.py stft mfcc) > $O/e7_race_pad16.txt 2>&1
grep -E "Warning|Error|SUMMARY|ok" $O/e7_race_pad16.txt | cut -c1-260 | head -20
(time timeout 600 compute-sanitizer --tool memcheck --kernel-regex kns=dsb200 --print-limit 30 python tools/sanitize_loop.py) > $O/e7_mem.txt 2>&1
grep -E "Error|SUMMARY|ok" $O/e7_mem.txt | cut -c1-260 | head -10
