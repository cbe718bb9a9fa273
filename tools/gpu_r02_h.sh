#!/bin/bash
# Where does the two-ring kernel lose time at batch 1?  Phase timelines of three builds / settings + column-halo parity & timing.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
PREV=$PWD/raft-tf_b200/lib/libraft_b200_prev.so
echo "== parity of the current build (row + column halo)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv2d or update_block or encoder or corr_pyramid" --timeout 300 --tb=line 2>&1 | tail -4
echo "== phase times: previous build"
RAFT_B200_LIB=$PREV timeout 200 python tools/phase_times.py 2>&1 | tail -11 | tee $O/phase_prev.log
echo "== phase times: new, halo on"
timeout 200 python tools/phase_times.py 2>&1 | tail -11 | tee $O/phase_new.log
echo "== phase times: new, halo off"
RAFT_B200_NO_ROWHALO=1 RAFT_B200_NO_COLHALO=1 timeout 200 python tools/phase_times.py 2>&1 | tail -11 | tee $O/phase_new_off.log
echo "== timings"
for w in update iterate; do
  echo -n "new row+col halo $w: "; timeout 200 python tools/micro.py $w 2>&1 | tail -1
  echo -n "new row halo only $w: "; RAFT_B200_NO_COLHALO=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1
  echo -n "previous build    $w: "; RAFT_B200_LIB=$PREV timeout 200 python tools/micro.py $w 2>&1 | tail -1
done | tee $O/colhalo_ab.log
echo "== e2e parity"
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 600 --tb=line 2>&1 | tail -4
