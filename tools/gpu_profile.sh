#!/bin/bash
# ncu pass (one GPU): launch list of one bench step + full captures of the headline kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RAFT_B200_NO_GRAPH=1   # ncu needs individual launches, not a graph replay
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "== full capture: lookup (micro benchmark)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_kernel -s 1 -c 2 -f -o gpurun_out/prof_lookup \
    python tools/micro.py lookup --reps 3 > gpurun_out/ncu_lookup.log 2>&1
echo "== full capture: update-step convs (micro benchmark)"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|conv_halo" -s 24 -c 11 -f -o gpurun_out/prof_update \
    python tools/micro.py update --reps 3 > gpurun_out/ncu_update.log 2>&1
ls -la gpurun_out | tail -8
