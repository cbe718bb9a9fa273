#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/diag_parity.py --tag cfg3-default 2>&1 | tail -4 | tee $O/diag_parity2.log
timeout 600 python tools/diag_parity.py --seed0 1004 --tag cfg3-seed1004 2>&1 | tail -4 | tee -a $O/diag_parity2.log
timeout 600 python tools/diag_parity.py --seed0 1005 --tag cfg3-seed1005 2>&1 | tail -4 | tee -a $O/diag_parity2.log
echo "== lookup A/B (v6 = optimised round-1 kernel, default; v5 = warp-per-pixel)"
for B in 1 8; do
  for fl in "" "--flush"; do
    echo -n "v6 B=$B $fl: "; RAFT_B200_LOOKUP_V4=1 timeout 200 python tools/micro.py lookup --B $B $fl 2>&1 | tail -1
    echo -n "v5 B=$B $fl: "; timeout 200 python tools/micro.py lookup --B $B $fl 2>&1 | tail -1
  done
done 2>&1 | tee $O/lookup_ab.log
RAFT_B200_LOOKUP_V4=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k lookup --timeout 200 --tb=short 2>&1 | tail -3
for w in iterate forward; do echo -n "v6 $w: "; RAFT_B200_LOOKUP_V4=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1; done | tee -a $O/lookup_ab.log
