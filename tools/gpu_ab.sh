#!/bin/bash
# same-box A/B: fused update-step kernel vs one kernel per conv, batch 1 / 8; kernel parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 2>&1 | tail -3
RAFT_B200_FUSED=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "update" --timeout 300 2>&1 | tail -2
for args in "" "--B 8"; do
  for rep in 1 2; do
  echo -n "per-conv $args : "; timeout 300 python tools/micro.py update $args 2>&1 | tail -1
  echo -n "fused    $args : "; RAFT_B200_FUSED=1 timeout 300 python tools/micro.py update $args 2>&1 | tail -1
  done
done
timeout 300 python tools/fused_times.py 2>&1 | tail -10 | awk '{print $1, "mma", $11-$9, "epi", $15-$13, "fence", $17-$15, "tail", $21-$19}' | tr '\n' ';'; echo
echo -n "bench per-conv: "; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
