#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -x --timeout 600 -k "encoder or e2e or forward or flow" 2>&1 | tail -2
for rep in 1 2; do
echo -n "bench fused stats : "; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
echo -n "bench stats pass  : "; RAFT_B200_NO_FUSED_STATS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
