#!/bin/bash
cd "$(dirname "$0")/.."
for env in "RAFT_B200_FH2_TRIGGER=1" "RAFT_B200_FH2_MMA=1"; do
  echo "== $env"; env $env timeout 300 python tools/phase_times.py 2>&1 | tail -11 | cut -c1-250
done
