"""Per-job timeline of the FUSED update-step kernel (RAFT_B200_FUSED=1; globaltimer stamps, update_fused.cu):
job start / grid barrier passed / first operands / MMA loop end / epilogue begin / stores issued / fence done / signalled."""
import os, sys
os.environ['RAFT_B200_FUSED'] = '1'
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
NC = 12
buf = torch.zeros(NC * 4096 * 8, dtype=torch.int64, device=dev)
lib.rb_debug_set_buffer(capi.ptr(buf))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step(); step()
lib.rb_debug_set_buffer(None)
g.replay(); torch.cuda.synchronize()
t = buf.view(NC, 4096, 8).cpu().double() / 1e3
names = ["convc1", "convf2", "convc2", "motion", "zr1", "q1", "zr2", "q2", "fh1", "fh2", "mask0", "mask2"]
t00 = t[0][:, 0][t[0][:, 0] > 0].min()
print("second step of a 2-step graph; us since the first CTA entered job 0; mean over CTAs (min..max for the end)")
for i in range(NC):
    m = t[i][:, 0] > 0
    if not m.any(): continue
    x = t[i][m] - t00
    f = lambda k: x[:, k][x[:, k] > -1e6].mean().item() if (x[:, k] > -1e6).any() else float('nan')
    print(f"{names[i]:7s} ctas={int(m.sum()):4d} start {f(0):6.1f} barrier {f(1):6.1f} first_ops {f(2):6.1f} mma_end {f(3):6.1f} "
          f"epi_begin {f(4):6.1f} stores {f(5):6.1f} fenced {f(6):6.1f} signalled {f(7):6.1f} (last {x[:,7].max():6.1f})")
