#!/bin/bash
# One gpurun call: parity tests (back ends in separate processes so a trap in one cannot poison the
# other), diagnostics, a short bench and the ncu launch list.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== simt/bit-exact tests"; timeout 900 python -m pytest tests -q -m gpu -k "not tc" -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_simt.log
echo "== diag"; timeout 300 python tools/diag_tc.py 2>&1 | tail -60 | tee gpurun_out/diag.log
echo "== tc tests"; timeout 900 python -m pytest tests -q -m gpu -k "tc" --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_tc.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
