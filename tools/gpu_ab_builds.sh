#!/bin/bash
# Same-box A/B of two builds: raft-tf_b200/lib/libraft_b200.so vs libraft_b200_prev.so (RAFT_B200_LIB override)
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 -k "update" 2>&1 | tail -1
for what in ${WHAT:-iterate}; do
for rep in 1 2; do
  echo -n "$what new : "; timeout 300 python tools/micro.py $what 2>&1 | tail -1
  echo -n "$what prev: "; RAFT_B200_LIB=$PWD/raft-tf_b200/lib/libraft_b200_prev.so timeout 300 python tools/micro.py $what 2>&1 | tail -1
done
done
