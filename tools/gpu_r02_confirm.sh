#!/bin/bash
# Last confirmation at HEAD: every -m gpu test file (one process each), smoke(), both bench arms.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
: > $O/pytest_gpu.log
for f in tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_variants.py tests/test_gpu_multi.py; do
  echo "-- $f" | tee -a $O/pytest_gpu.log
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 --maxfail=3 --tb=short 2>&1 | grep -vE "^\s*$" | tail -12 | tee -a $O/pytest_gpu.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.log
timeout 900 python bench.py 2>$O/bench_err.log | tail -1 > $O/bench_default.json; cut -c1-700 $O/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>$O/bench_err.log | tail -1 > $O/bench_reference.json; cut -c1-300 $O/bench_reference.json
