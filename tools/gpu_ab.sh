#!/bin/bash
# same-box A/B: fused update-step kernel vs one kernel per conv; then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
RAFT_B200_FUSED=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "update" --timeout 300 2>&1 | tail -3
for rep in 1 2; do
  echo -n "per-conv kernels : "; timeout 300 python tools/micro.py update 2>&1 | tail -1
  echo -n "fused            : "; RAFT_B200_FUSED=1 timeout 300 python tools/micro.py update 2>&1 | tail -1
done
timeout 300 python tools/fused_times.py 2>&1 | tail -11
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -4
echo -n "bench fused   : "; RAFT_B200_FUSED=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
