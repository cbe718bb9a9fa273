#!/bin/bash
# Same-box A/B template: one knob on/off on the micro benchmarks (update step, lookup + update step), B = 1 and 8.
#   usage: tools/gpu_ab.sh RAFT_B200_NO_PDL      (any environment knob of README.md)
cd "$(dirname "$0")/.."
knob=${1:-RAFT_B200_NO_PDL}
for args in "" "--B 8"; do
  for what in update iterate; do
    echo -n "default      $what $args : "; timeout 300 python tools/micro.py $what $args 2>&1 | tail -1
    echo -n "$knob=1 $what $args : "; env $knob=1 timeout 300 python tools/micro.py $what $args 2>&1 | tail -1
  done
done
