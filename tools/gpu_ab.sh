#!/bin/bash
cd "$(dirname "$0")/.."
for lim in 0 38 28 55; do
  echo -n "convf2 ctas=$lim iterate : "; RAFT_B200_CONVF2_CTAS=$lim timeout 300 python tools/micro.py iterate 2>&1 | tail -1
done
echo -n "convf2 ctas=0 iterate : "; RAFT_B200_CONVF2_CTAS=0 timeout 300 python tools/micro.py iterate 2>&1 | tail -1
