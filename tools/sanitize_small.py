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
