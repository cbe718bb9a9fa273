import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
import torch
from raft_b200 import synth
from raft_b200.engine import RaftEngine
dev = torch.device("cuda:0")
p = synth.make_weights(False)
l, r = synth.make_batch(1, 64, 96)
l, r = torch.from_numpy(l).to(dev), torch.from_numpy(r).to(dev)
def snap(e):
    return {k: getattr(e, k).clone() for k in ("fmaps", "cmap", "pyramid", "coords1", "flow_up")}
e1 = RaftEngine(p, iters=3, device=dev, use_graph=False)
e1.forward(l, r); torch.cuda.synchronize(); s1 = snap(e1)
e1.forward(l, r); torch.cuda.synchronize(); s2 = snap(e1)
e2 = RaftEngine(p, iters=3, device=dev, use_graph=False)
e2.forward(l, r); torch.cuda.synchronize(); s3 = snap(e2)
for k in s1:
    print(k, "run1 vs run2 (same engine):", (s1[k] - s2[k]).abs().max().item(), " run1 vs fresh engine:", (s1[k] - s3[k]).abs().max().item())
e3 = RaftEngine(p, iters=3, device=dev, use_graph=True)
e3.forward(l, r); torch.cuda.synchronize(); s4 = snap(e3)
e3.forward(l, r); torch.cuda.synchronize(); s5 = snap(e3)
for k in s1:
    print(k, "eager vs graph:", (s1[k] - s4[k]).abs().max().item(), " graph replay 1 vs 2:", (s4[k] - s5[k]).abs().max().item())
