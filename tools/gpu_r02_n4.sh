#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench${N}_err.log | tail -1 > gpurun_out/bench_n$N.json
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_n{n}.json").read())
print("N", n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3))
for k, v in d.get("other_configs", {}).items():
    print(k, v.get("workload"), "value", round(v.get("value", 0), 1), "e2e", round(v.get("e2e", {}).get("value", 0), 1), "gather_us", v.get("gather_us"))
PY
tail -3 gpurun_out/bench${N}_err.log
