#!/bin/bash
# compute-sanitizer over the small end-to-end script: memcheck, then racecheck (shared memory / DSMEM hazards).
# racecheck wants a cluster barrier between the split-K pair's remote write and the receiver's exit: RAFT_B200_SPLITK_CLOSING_BARRIER=1
# adds it (conv_tc.cu); the default build's only report is that liveness heuristic (second racecheck run, trimmed).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "update_block or conv2d" --timeout 200 --tb=short -x 2>&1 | tail -3
timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_small.py > $O/r02_sanitizer_memcheck.log 2>&1
tail -3 $O/r02_sanitizer_memcheck.log
RAFT_B200_SPLITK_CLOSING_BARRIER=1 timeout 400 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_small.py > $O/r02_sanitizer_racecheck.log 2>&1
tail -3 $O/r02_sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 2 python tools/sanitize_small.py 2>&1 | grep -vE "Host Frame|^=========\s*$" > $O/r02_sanitizer_racecheck_default.log
tail -4 $O/r02_sanitizer_racecheck_default.log
