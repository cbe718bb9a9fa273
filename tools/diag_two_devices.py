"""Two engines on two devices in one process, step by step (run on a >= 2-GPU box; CUDA_LAUNCH_BLOCKING=1 localises faults)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
from types import SimpleNamespace
import torch
from raft_b200 import capi, synth
from networks.RAFT import RAFT
lib = capi.lib
print("devices", torch.cuda.device_count(), flush=True)
for d in (0, 1, 0):
    with torch.cuda.device(d):
        g = torch.empty(2 * 5 * 7 * 2, device=f"cuda:{d}")
        capi.check(lib.rb_coords_grid(capi.ptr(g), 2, 5, 7, capi.stream()))
        torch.cuda.synchronize(d)
        print("coords_grid on", d, "ok", float(g.sum()), flush=True)
p = synth.make_weights(False)
l, r = synth.make_batch(1, 96, 160)
outs = []
for d in ("cuda:0", "cuda:1", "cuda:0", "cuda:1"):
    print("engine on", d, flush=True)
    m = RAFT((96, 160, 3), SimpleNamespace(small=False), iters=4, device=d).load(p)
    o = m.forward(l, r)
    for i in range(torch.cuda.device_count()):
        torch.cuda.synchronize(i)
    print("  forward ok", o.device, float(o.abs().max()), flush=True)
    outs.append(o.cpu())
print("equal:", [bool(torch.equal(outs[0], x)) for x in outs], flush=True)
