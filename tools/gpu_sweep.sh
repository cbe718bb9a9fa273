#!/bin/bash
# stage-count sweep of the per-tap conv kernel on the update block (micro benchmark, graph-timed)
cd "$(dirname "$0")/.."
for st in 1 2 3 4 5 6; do
  echo "== NO_HALO stages<=$st"; RAFT_B200_NO_HALO=1 RAFT_B200_TC_STAGES=$st python tools/micro.py update 2>&1 | tail -1
done
echo "== HALO"; python tools/micro.py update 2>&1 | tail -1
