#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2 3; do
  echo -n "halo all      : "; python tools/micro.py update 2>&1 | tail -1
  echo -n "halo N>=96    : "; RAFT_B200_HALO_MIN_N=96 python tools/micro.py update 2>&1 | tail -1
  echo -n "halo N>=64    : "; RAFT_B200_HALO_MIN_N=64 python tools/micro.py update 2>&1 | tail -1
  echo -n "no halo       : "; RAFT_B200_NO_HALO=1 python tools/micro.py update 2>&1 | tail -1
done
