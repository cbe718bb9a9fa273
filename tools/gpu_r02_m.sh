#!/bin/bash
# convf1 (7x7x2) on the tensor cores through the 8-pixel window view: parity + same-box A/B against the CUDA-core kernel.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "update_block" --timeout 300 --tb=short -x 2>&1 | tail -8 | tee $O/convf1_parity.log
for r in 1 2; do
  for B in 1 4 8; do
    echo -n "tensor-core convf1 B=$B update: "; timeout 200 python tools/micro.py update --B $B 2>&1 | tail -1
    echo -n "CUDA-core  convf1 B=$B update: "; RAFT_B200_CONVF1_SIMT=1 timeout 200 python tools/micro.py update --B $B 2>&1 | tail -1
  done
  echo -n "tensor-core convf1 iterate: "; timeout 200 python tools/micro.py iterate 2>&1 | tail -1
  echo -n "tensor-core, no CTA limit iterate: "; RAFT_B200_CONVF1_CTAS=0 timeout 200 python tools/micro.py iterate 2>&1 | tail -1
  echo -n "CUDA-core  convf1 iterate: "; RAFT_B200_CONVF1_SIMT=1 timeout 200 python tools/micro.py iterate 2>&1 | tail -1
done | tee $O/convf1_ab.log
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -q -m gpu --timeout 600 --tb=short 2>&1 | tail -6 | tee $O/convf1_e2e.log
