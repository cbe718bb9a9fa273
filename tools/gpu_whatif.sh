#!/bin/bash
# timing experiments on the fused update-step kernel: which resource bounds the MMA loop?
cd "$(dirname "$0")/.."
for wi in 0 1 2 3 4 8 10 14 16 17 32 48 62; do
  echo -n "whatif=$wi : "; RAFT_B200_FUSED=1 RAFT_B200_WHATIF=$wi timeout 300 python tools/micro.py update 2>&1 | tail -1
done
