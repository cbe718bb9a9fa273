#!/bin/bash
# Closing cluster barrier of the split-K conv (early arrive): parity, cost against the build without it, racecheck.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
PREV=$PWD/raft-tf_b200/lib/libraft_b200_prev.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "update_block or conv2d" --timeout 300 --tb=short -x 2>&1 | tail -4
for r in 1 2 3; do
  for w in update iterate; do
    echo -n "closing barrier $w: "; timeout 200 python tools/micro.py $w 2>&1 | tail -1
    echo -n "without         $w: "; RAFT_B200_LIB=$PREV timeout 200 python tools/micro.py $w 2>&1 | tail -1
  done
done | tee $O/splitk_barrier_ab.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_small.py > $O/r02_sanitizer_memcheck.log 2>&1; tail -3 $O/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_small.py > $O/r02_sanitizer_racecheck.log 2>&1; tail -3 $O/r02_sanitizer_racecheck.log
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 600 --tb=line 2>&1 | tail -3
