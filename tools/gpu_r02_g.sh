#!/bin/bash
# Row halo validated: full parity suite, then same-box A/B of builds (previous kernel / new kernel with and without row halo).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
: > $O/pytest_gpu.log
for f in tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_variants.py; do
  echo "-- $f" | tee -a $O/pytest_gpu.log
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 --maxfail=3 --tb=line -s 2>&1 | grep -vE "^\s*$" | tail -14 | tee -a $O/pytest_gpu.log
done
echo "== same-box A/B of builds"
PREV=$PWD/raft-tf_b200/lib/libraft_b200_prev.so
for rep in 1 2; do
  for w in update iterate encoder forward; do
    echo -n "new (row halo) $w: "; timeout 200 python tools/micro.py $w 2>&1 | tail -1
    echo -n "new, halo off  $w: "; RAFT_B200_NO_ROWHALO=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1
    echo -n "previous build $w: "; RAFT_B200_LIB=$PREV timeout 200 python tools/micro.py $w 2>&1 | tail -1
  done
done | tee $O/rowhalo_ab2.log
for B in 4 8; do
  echo -n "new (row halo) B=$B iterate: "; timeout 200 python tools/micro.py iterate --B $B 2>&1 | tail -1
  echo -n "previous build B=$B iterate: "; RAFT_B200_LIB=$PREV timeout 200 python tools/micro.py iterate --B $B 2>&1 | tail -1
done | tee -a $O/rowhalo_ab2.log
echo "== bench"
timeout 900 python bench.py 2>$O/bench_err.log | tail -1 | tee $O/bench_default.json | cut -c1-700
