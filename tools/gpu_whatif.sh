#!/bin/bash
# timing experiments on the fused update-step kernel: smem ring depth
cd "$(dirname "$0")/.."
for st in 3 2 1; do
  echo "stages=$st : "; RAFT_B200_FUSED=1 RAFT_B200_FUSED_STAGES=$st timeout 300 python tools/micro.py update 2>&1 | tail -1
  RAFT_B200_FUSED_STAGES=$st timeout 300 python tools/fused_times.py 2>&1 | tail -10 | awk '{print $1, "mma", $11-$9, "epi", $15-$13, "fence", $17-$15, "tail", $21-$19}' | tr '\n' ';'; echo
done
