#!/bin/bash
# ncu --set full of the update-step convs and the lookup kernel (micro benchmarks, one GPU)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RAFT_B200_NO_PDL=1   # ncu serialises kernels anyway; keeps the per-kernel durations free of PDL overlap
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|flow_conv7" -s 22 -c 11 -f -o gpurun_out/prof_update \
    python tools/micro.py update --reps 2 > gpurun_out/ncu_update.log 2>&1
tail -2 gpurun_out/ncu_update.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_kernel -s 1 -c 2 -f -o gpurun_out/prof_lookup \
    python tools/micro.py lookup --reps 3 > gpurun_out/ncu_lookup.log 2>&1
tail -2 gpurun_out/ncu_lookup.log
ls -la gpurun_out/*.ncu-rep
