#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/test_gpu_variants.py -q -x --timeout 1500 2>&1 | tail -2
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['us_per_launch_group'], d['roofline_lookup']['frac'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
bash tools/batch_sweep.sh 2>&1 | tee gpurun_out/batch_sweep.log
