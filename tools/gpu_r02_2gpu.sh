#!/bin/bash
# Round 2, two-GPU pass: NCCL batch shard == single GPU (bit exact), two devices in one process, bench at N=2 with the
# other BASELINE configs (all_gather inside the e2e span); plus same-box A/B repeats on GPU 0.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
echo "== lookup kernels after the 10-row change"
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "lookup or volume_free" --timeout 200 --tb=short 2>&1 | tail -3
echo "== multi-GPU tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 800 --tb=short -s 2>&1 | grep -vE "^\s*$" | tail -15 | tee $O/pytest_multi.log
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
echo "== same-box A/B (GPU 0): conv2 fold on/off, 3 repeats"
export CUDA_VISIBLE_DEVICES=0
for rep in 1 2 3; do
  for B in 1 8; do
    echo -n "fold on  B=$B: "; timeout 200 python tools/micro.py iterate --B $B 2>&1 | tail -1
    echo -n "fold off B=$B: "; RAFT_B200_NO_FH2_FUSE=1 timeout 200 python tools/micro.py iterate --B $B 2>&1 | tail -1
  done
done | tee $O/ab_fh2_r02.log
for B in 1 8; do for fl in "" "--flush"; do echo -n "lookup B=$B $fl: "; timeout 200 python tools/micro.py lookup --B $B $fl 2>&1 | tail -1; done; done | tee $O/lookup_10rows.log
echo -n "corr: "; timeout 200 python tools/micro.py corr 2>&1 | tail -1
echo -n "forward: "; timeout 200 python tools/micro.py forward 2>&1 | tail -1
