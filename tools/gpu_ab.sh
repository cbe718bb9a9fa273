#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2 3; do
  echo -n "new (prefetch)   : "; python tools/micro.py update 2>&1 | tail -1
  echo -n "new (no prefetch): "; RAFT_B200_NO_EPI_PREFETCH=1 python tools/micro.py update 2>&1 | tail -1
  echo -n "prev build       : "; RAFT_B200_LIB=$PWD/raft-tf_b200/lib/libraft_b200_prev.so python tools/micro.py update 2>&1 | tail -1
done
