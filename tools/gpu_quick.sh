#!/bin/bash
# Fast iteration loop: full GPU parity suite + one bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -3 | tee gpurun_out/bench.log
