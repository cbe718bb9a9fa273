"""Run one of the BASELINE.json configs on the local GPU and print pairs/s (CUDA events, L2 flushed between steps).
usage: python tools/run_config.py {c2|c3|c4_per_gpu|c5_per_gpu} [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
from types import SimpleNamespace
import numpy as np, torch
from raft_b200 import synth
from networks.RAFT import RAFT
CFG = {"c2": (False, 1, 436, 1024, 32), "c3": (False, 8, 540, 960, 32), "c4_per_gpu": (False, 4, 436, 1024, 32),
       "c5_per_gpu": (True, 8, 768, 1024, 20)}
small, B, H, W, iters = CFG[sys.argv[1]]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
m = RAFT((H, W, 3), SimpleNamespace(small=small), iters=iters, batch=B, device=dev).load(synth.make_weights(small))
one = synth.make_batch(1, H, W)
l = torch.from_numpy(np.repeat(one[0], B, 0)).to(dev); r = torch.from_numpy(np.repeat(one[1], B, 0)).to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for _ in range(3): out = m.forward(l, r)
torch.cuda.synchronize()
assert torch.isfinite(out).all()
assert all(torch.equal(out[0], out[i]) for i in range(1, B)), "identical samples must give identical flows"
ts = []
for _ in range(steps):
    flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m.forward(l, r); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
t = sum(ts) / len(ts)
print(f"{sys.argv[1]}: small={small} B={B} {H}x{W} iters={iters}: {t:.2f} ms/step, {B / t * 1e3:.1f} pairs/s, "
      f"max|flow|={out.abs().max().item():.2f}, mem={torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
