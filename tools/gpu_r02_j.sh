#!/bin/bash
# Split-K of the flow head's conv2: parity, bit-equality of the cluster and single-CTA executions, timing.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== update block / conv parity"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "update_block or conv2d or raft_iterate" --timeout 300 --tb=short 2>&1 | tail -15 | tee $O/splitk_parity.log
echo "== cluster vs single-CTA two-halves: bit-equal?"
timeout 300 python - <<'PY' 2>&1 | tail -5 | tee $O/splitk_equal.log
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, "raft-tf_b200")
from types import SimpleNamespace
from raft_b200 import synth
from networks.RAFT import RAFT
p = synth.make_weights(False)
l, r = synth.make_batch(1, 440, 1024)
m = RAFT((440, 1024, 3), SimpleNamespace(small=False), iters=6, batch=1).load(p)
out = m.forward(l, r)
torch.save(out.cpu(), sys.argv[1])
'''
subprocess.run([sys.executable, "-c", code, "gpurun_out/sk_a.pt"], check=True)
subprocess.run([sys.executable, "-c", code, "gpurun_out/sk_b.pt"], check=True, env=dict(os.environ, RAFT_B200_NO_SPLITK_CLUSTER="1"))
subprocess.run([sys.executable, "-c", code, "gpurun_out/sk_c.pt"], check=True, env=dict(os.environ, RAFT_B200_NO_SPLITK="1"))
import torch
a, b, c = (torch.load(f"gpurun_out/sk_{x}.pt") for x in "abc")
print("cluster == single-CTA halves:", torch.equal(a, b), " max |halves - one accumulator|:", (a - c).abs().max().item())
os.remove("gpurun_out/sk_a.pt"); os.remove("gpurun_out/sk_b.pt"); os.remove("gpurun_out/sk_c.pt")
PY
echo "== timings (ABAB)"
for r in 1 2; do
  for w in update iterate; do
    echo -n "split-K cluster   $w: "; timeout 200 python tools/micro.py $w 2>&1 | tail -1
    echo -n "halves on one CTA $w: "; RAFT_B200_NO_SPLITK_CLUSTER=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1
    echo -n "no split-K        $w: "; RAFT_B200_NO_SPLITK=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1
  done
done | tee $O/splitk_ab.log
echo "== full-size / e2e / configs"
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_e2e.py tests/test_gpu_configs.py -q -m gpu --timeout 600 --tb=line 2>&1 | tail -5 | tee $O/splitk_e2e.log
echo "== bench"
timeout 600 python bench.py --headline-only 2>$O/bench_err.log | tail -1 | tee $O/bench_splitk.json | cut -c1-400
