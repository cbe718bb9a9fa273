#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2; do
  echo -n "late trigger  : "; python tools/micro.py update 2>&1 | tail -1
  echo -n "early trigger : "; RAFT_B200_PDL_EARLY=1 python tools/micro.py update 2>&1 | tail -1
done
