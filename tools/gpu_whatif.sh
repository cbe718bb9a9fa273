#!/bin/bash
# timing experiments on the fused update-step kernel (after the issue-loop fixes): what bounds the k-iteration now?
cd "$(dirname "$0")/.."
for wi in 0 1 16 10 14 30; do
  echo "whatif=$wi : "; RAFT_B200_FUSED=1 RAFT_B200_WHATIF=$wi timeout 300 python tools/micro.py update 2>&1 | tail -1
  RAFT_B200_WHATIF=$wi timeout 300 python tools/fused_times.py 2>&1 | tail -10 | awk '{print $1, "mma", $11-$9}' | tr '\n' ';'; echo
done
