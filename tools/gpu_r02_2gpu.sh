#!/bin/bash
# Round 2, two-GPU pass: two devices in one process (rb_set_device), NCCL batch shard == single GPU (bit exact),
# bench at N=2 with the other BASELINE configs (all_gather inside the e2e span), reference arm under torchrun.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
echo "== multi-GPU tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 800 --tb=short -s 2>&1 | grep -vE "^\s*$" | tail -25 | tee $O/pytest_multi.log
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2>$O/bench2_err.log | tail -1 | tee $O/bench_n2.json | cut -c1-400
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2.json").read())
    print("N=2 value", d["value"], "e2e", d["e2e"]["value"])
    for k, v in d.get("other_configs", {}).items():
        print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "e2e"}, "e2e", v.get("e2e", {}).get("value"))
except Exception as e:
    print("bench_n2 parse failed", e)
PY
tail -5 $O/bench2_err.log
echo "== reference arm under torchrun (rank 0 works, rank 1 exits)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>>$O/bench2_err.log | tail -1 | cut -c1-300
