#!/bin/bash
# compute-sanitizer over the small end-to-end script (memcheck, then racecheck on shared memory / DSMEM hazards)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
tail -12 gpurun_out/r02_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
tail -12 gpurun_out/r02_sanitizer_racecheck.log
