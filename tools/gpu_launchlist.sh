#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RAFT_B200_NO_GRAPH=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
