#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== two devices, graphs on"; timeout 300 python tools/diag_two_devices.py 2>&1 | tail -14
echo "== multi-GPU tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 800 --tb=short 2>&1 | grep -vE "^\s*$" | tail -8 | tee gpurun_out/pytest_multi.log
