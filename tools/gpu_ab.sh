#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -x --timeout 600 2>&1 | tail -2
for rep in 1 2; do
  echo -n "iterate x4 default : "; timeout 300 python tools/micro.py iterate 2>&1 | tail -1
  echo -n "iterate x4 prev lib: "; RAFT_B200_LIB=$PWD/raft-tf_b200/lib/libraft_b200_prev.so timeout 300 python tools/micro.py iterate 2>&1 | tail -1
done
echo -n "bench default: "; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], d.get('roofline_lookup'))"
bash tools/gpu_launchlist.sh
