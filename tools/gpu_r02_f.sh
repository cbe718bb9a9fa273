#!/bin/bash
# Row-halo experiment: conv parity with / without the descriptor base offset, then timings on/off.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== conv parity, default (base offset = tap)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv2d or update_block or encoder or corr_pyramid" --timeout 300 --tb=line 2>&1 | tail -6 | tee $O/rowhalo_parity.log
echo "== conv parity, RAFT_B200_DESC_NOBOFF=1"
RAFT_B200_DESC_NOBOFF=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv2d or update_block or encoder" --timeout 300 --tb=line 2>&1 | tail -6 | tee -a $O/rowhalo_parity.log
echo "== conv parity, RAFT_B200_NO_ROWHALO=1"
RAFT_B200_NO_ROWHALO=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv2d or update_block or encoder or corr_pyramid" --timeout 300 --tb=line 2>&1 | tail -4 | tee -a $O/rowhalo_parity.log
echo "== timings (ABAB)"
for rep in 1 2; do
  for w in update iterate encoder corr forward; do
    echo -n "rowhalo on  $w: "; timeout 200 python tools/micro.py $w 2>&1 | tail -1
    echo -n "rowhalo off $w: "; RAFT_B200_NO_ROWHALO=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1
  done
done | tee $O/rowhalo_ab.log
echo -n "rowhalo on  B=8 iterate: "; timeout 200 python tools/micro.py iterate --B 8 2>&1 | tail -1 | tee -a $O/rowhalo_ab.log
echo -n "rowhalo off B=8 iterate: "; RAFT_B200_NO_ROWHALO=1 timeout 200 python tools/micro.py iterate --B 8 2>&1 | tail -1 | tee -a $O/rowhalo_ab.log
echo "== e2e tests"
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 600 --tb=line 2>&1 | tail -5
