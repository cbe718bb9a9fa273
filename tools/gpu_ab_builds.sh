#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -x --timeout 600 2>&1 | tail -2
for rep in 1 2; do
  echo -n "iterate x4 new : "; timeout 300 python tools/micro.py iterate 2>&1 | tail -1
  echo -n "iterate x4 prev: "; RAFT_B200_LIB=$PWD/raft-tf_b200/lib/libraft_b200_prev.so timeout 300 python tools/micro.py iterate 2>&1 | tail -1
done
echo -n "B=8 new : "; timeout 300 python tools/micro.py iterate --B 8 2>&1 | tail -1
echo -n "B=8 prev: "; RAFT_B200_LIB=$PWD/raft-tf_b200/lib/libraft_b200_prev.so timeout 300 python tools/micro.py iterate --B 8 2>&1 | tail -1
