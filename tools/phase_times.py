"""Per-phase timeline of the tensor-core convs of one update step (globaltimer stamps recorded by the kernels):
launch->prologue end, PDL wait, first operands landed, MMA loop, epilogue, teardown; mean over CTAs, in us."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
import torch
from raft_b200 import capi, synth
from raft_b200.weights import pack_update_block
dev = torch.device("cuda:0"); lib = capi.lib
B, h, w, s = 1, 55, 128, 0
pyr = torch.randn(capi.size_query(lib.rb_corr_pyramid_bytes, B, h, w) // 4, device=dev)
grid = torch.stack(torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy"), -1).float()[None]
coords = (grid + torch.rand(B, h, w, 2) * 8 - 4).to(dev).contiguous()
ws = torch.zeros(capi.size_query(lib.rb_update_workspace_bytes, s, B, h, w), dtype=torch.uint8, device=dev)
blob = pack_update_block(synth.make_weights(False), False, dev)
net = torch.tanh(torch.randn(B, h, w, 128, device=dev)); inp = torch.relu(torch.randn(B, h, w, 128, device=dev))
capi.check(lib.rb_update_set_state(s, capi.ptr(blob), capi.ptr(ws), capi.ptr(net), capi.ptr(inp), B, h, w, capi.stream()))
capi.check(lib.rb_update_lookup(s, capi.ptr(ws), capi.ptr(pyr), capi.ptr(coords), B, h, w, capi.stream()))
c1 = coords.clone()
step = lambda: capi.check(lib.rb_update_step(s, capi.ptr(blob), capi.ptr(ws), capi.ptr(c1), None, None, B, h, w, capi.stream()))
for _ in range(3): step()
torch.cuda.synchronize()
NC = 13
buf = torch.zeros(NC * 4096 * 8, dtype=torch.int64, device=dev)
lib.rb_debug_set_buffer(capi.ptr(buf))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step(); step()
lib.rb_debug_set_buffer(None)
g.replay(); torch.cuda.synchronize()
t = buf.view(NC, 4096, 8).cpu().double()
names = ["convf1(side)", "convf2(side)", "convc1", "convc2", "motion", "zr1", "q1", "zr2", "q2", "fh1", "fh2"]
print("second step of a 2-step graph; us relative to each conv's earliest CTA start")
t0_prev = None
for i in range(NC):
    m = t[i][:, 0] > 0
    if not m.any(): continue
    x = t[i][m] / 1e3
    base = x[:, 0].min()
    f = lambda k: (x[:, k].mean() - base).item()
    print(f"{names[i] if i < len(names) else i:13s} ctas={int(m.sum()):4d} start+{x[:,0].max()-base:5.1f} prologue_end {f(1):5.1f} pdl_wait_end {f(2):5.1f} first_ops {f(3):5.1f} "
          f"mma_end {f(4):5.1f} epi_begin {f(5):5.1f} epi_end {f(6):5.1f} cta_end {f(7):5.1f} | last cta_end {x[:,7].max()-base:5.1f}"
          + (f" | gap from prev conv end {base - t0_prev:5.1f}" if t0_prev is not None else ""))
    t0_prev = x[:, 7].max()
