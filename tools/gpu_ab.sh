#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -x --timeout 600 2>&1 | tail -3
for rep in 1 2; do
  echo -n "iterate x4 fused flow head : "; timeout 300 python tools/micro.py iterate 2>&1 | tail -1
  echo -n "iterate x4 two convs       : "; RAFT_B200_NO_FH_FUSE=1 timeout 300 python tools/micro.py iterate 2>&1 | tail -1
done
echo -n "B=8 fused : "; timeout 300 python tools/micro.py iterate --B 8 2>&1 | tail -1
echo -n "B=8 two   : "; RAFT_B200_NO_FH_FUSE=1 timeout 300 python tools/micro.py iterate --B 8 2>&1 | tail -1
