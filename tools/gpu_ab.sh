#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -x --timeout 600 -k "update or e2e or flow or forward" 2>&1 | tail -2
for seg in 32 16 8; do
  echo -n "SEG=$seg iterate      : "; RAFT_B200_CONV7_SEG=$seg timeout 300 python tools/micro.py iterate 2>&1 | tail -1
  echo -n "SEG=$seg update       : "; RAFT_B200_CONV7_SEG=$seg timeout 300 python tools/micro.py update 2>&1 | tail -1
done
echo -n "SEG=32 iterate B8 : "; RAFT_B200_CONV7_SEG=32 timeout 300 python tools/micro.py iterate --B 8 2>&1 | tail -1
echo -n "SEG=8 iterate B8  : "; RAFT_B200_CONV7_SEG=8 timeout 300 python tools/micro.py iterate --B 8 2>&1 | tail -1
