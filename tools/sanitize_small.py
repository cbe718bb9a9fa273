"""Tiny end-to-end + kernel calls for compute-sanitizer (memcheck / racecheck): both variants, 64x96 frames, 2 iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
from types import SimpleNamespace
import torch
from raft_b200 import synth
from networks.RAFT import RAFT
for small in (False, True):
    p = synth.make_weights(small)
    l, r = synth.make_batch(1, 64, 96)
    m = RAFT((64, 96, 3), SimpleNamespace(small=small), iters=2, device="cuda:0").load(p)
    m.engine().use_graph = False
    f = m.forward(l, r)
    torch.cuda.synchronize()
    print("small", small, "flow", tuple(f.shape), float(f.abs().max()))
# round 2: strided / overlapping-window encoder views at odd sizes (TF SAME pads 3|3, 1|1), volume-free lookup, batch 2
from raft_b200.encoders import CudaEncoder
for small, name, norm, od in ((False, "cnet", "batch", 256), (True, "fnet", "instance", 128)):
    enc = CudaEncoder(synth.make_weights(small), name, small, norm, od, torch.device("cuda:0"))
    out = enc(torch.rand(2, 71, 99, 3, device="cuda:0"))
    torch.cuda.synchronize()
    print("encoder", name, "odd size ->", tuple(out.shape), float(out.abs().max()))
p = synth.make_weights(False)
l, r = synth.make_batch(2, 64, 96)
m = RAFT((64, 96, 3), SimpleNamespace(small=False), iters=2, batch=2, device="cuda:0", volume_free=True).load(p)
m.engine().use_graph = False
f = m.forward(l, r)
torch.cuda.synchronize()
print("volume-free B=2 flow", tuple(f.shape), float(f.abs().max()))
