#!/bin/bash
# Round-end validation on one GPU: full GPU suite, smoke(), default bench line, batch sweep, the other BASELINE configs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 1200 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-400
bash tools/batch_sweep.sh 2>&1 | tee gpurun_out/batch_sweep.log
for c in c3 c4_per_gpu c5_per_gpu; do timeout 300 python tools/run_config.py $c 3 2>&1 | tail -1; done | tee gpurun_out/configs.log
