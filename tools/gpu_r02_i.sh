#!/bin/bash
# Encoders without gather passes: strided TMA views, space-to-depth stem (overlapping view), residual in the epilogue.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
PREV=$PWD/raft-tf_b200/lib/libraft_b200_prev.so
echo "== encoder / conv parity (tc + simt)"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "encoder or conv2d" --timeout 300 --tb=short 2>&1 | tail -30 | tee $O/enc_parity.log
echo "== same with materialised stem windows"
RAFT_B200_STEM_WINDOWS=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "encoder and tc" --timeout 300 --tb=line 2>&1 | tail -6 | tee $O/enc_parity_windows.log
echo "== timings (ABAB)"
for r in 1 2; do
  echo -n "new               encoder: "; timeout 200 python tools/micro.py encoder 2>&1 | tail -1
  echo -n "new, windows      encoder: "; RAFT_B200_STEM_WINDOWS=1 timeout 200 python tools/micro.py encoder 2>&1 | tail -1
  echo -n "previous build    encoder: "; RAFT_B200_LIB=$PREV timeout 200 python tools/micro.py encoder 2>&1 | tail -1
  echo -n "new               forward: "; timeout 200 python tools/micro.py forward 2>&1 | tail -1
  echo -n "previous build    forward: "; RAFT_B200_LIB=$PREV timeout 200 python tools/micro.py forward 2>&1 | tail -1
done | tee $O/enc_ab.log
echo "== e2e parity"
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 600 --tb=line 2>&1 | tail -4 | tee $O/enc_e2e.log
echo "== launch list of the encoders"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/enc_launches.csv python tools/micro.py encoder --n 1 --reps 1 > $O/enc_ncu.log 2>&1
tail -2 $O/enc_ncu.log
