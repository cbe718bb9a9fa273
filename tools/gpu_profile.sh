#!/bin/bash
# ncu pass (one GPU): launch list of one bench step + full captures of the two headline kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RAFT_B200_NO_GRAPH=1   # ncu needs individual launches, not a graph replay
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log
echo "== full capture: lookup"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_kernel -s 40 -c 2 -f -o gpurun_out/prof_lookup \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lookup.log 2>&1
echo "== full capture: conv_tc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 400 -c 12 -f -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out
