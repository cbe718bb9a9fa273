#!/bin/bash
# ncu --set full on one micro benchmark:  gpu_prof_one.sh <what> <kernel regex> <skip> <count> [extra micro args]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what=$1; regex=$2; skip=$3; cnt=$4; shift 4
python tools/micro.py $what "$@" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $cnt -f -o gpurun_out/prof_$what \
   python tools/micro.py $what --reps 3 "$@" > gpurun_out/ncu_$what.log 2>&1
tail -3 gpurun_out/ncu_$what.log
