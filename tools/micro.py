"""Micro-benchmark / profiling driver for single kernels (run under gpurun, optionally under ncu).
usage: python tools/micro.py lookup|update|iterate|corr|encoder|forward [--B 1] [--reps 5] [--flush]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
import torch
from raft_b200 import capi, synth
from raft_b200.weights import pack_update_block

ap = argparse.ArgumentParser()
ap.add_argument("what")
ap.add_argument("--B", type=int, default=1)
ap.add_argument("--h", type=int, default=55)
ap.add_argument("--w", type=int, default=128)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--small", action="store_true")
ap.add_argument("--flush", action="store_true", help="flush L2 (256 MiB write) before every call; time = graph with - graph without")
ap.add_argument("--n", type=int, default=20, help="calls per timed graph")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib, B, h, w, s = capi.lib, a.B, a.h, a.w, int(a.small)
r = 3 if a.small else 4
g = torch.Generator(device="cpu").manual_seed(1234)
pyr_bytes = capi.size_query(lib.rb_corr_pyramid_bytes, B, h, w)
pyr = torch.randn(pyr_bytes // 4, device=dev)
grid = torch.stack(torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy"), -1).float()[None].repeat(B, 1, 1, 1)
coords = (grid + (torch.rand(B, h, w, 2, generator=g) * 16 - 8)).to(dev).contiguous()
wsb = capi.size_query(lib.rb_update_workspace_bytes, s, B, h, w)
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
st = None
if a.what == "lookup":
    fn = lambda: capi.check(lib.rb_update_lookup(s, capi.ptr(ws), capi.ptr(pyr), capi.ptr(coords), B, h, w, capi.stream()))
elif a.what == "update":
    blob = pack_update_block(synth.make_weights(a.small), a.small, dev)
    hid, ctx = (96, 64) if a.small else (128, 128)
    net = torch.tanh(torch.randn(B, h, w, hid, device=dev)); inp = torch.relu(torch.randn(B, h, w, ctx, device=dev))
    capi.check(lib.rb_update_set_state(s, capi.ptr(blob), capi.ptr(ws), capi.ptr(net), capi.ptr(inp), B, h, w, capi.stream()))
    capi.check(lib.rb_update_lookup(s, capi.ptr(ws), capi.ptr(pyr), capi.ptr(coords), B, h, w, capi.stream()))
    c1 = coords.clone()
    fn = lambda: capi.check(lib.rb_update_step(s, capi.ptr(blob), capi.ptr(ws), capi.ptr(c1), None, None, B, h, w, capi.stream()))
elif a.what == "iterate":  # lookup + update step, as inside rb_raft_iterate (4 iterations per call)
    blob = pack_update_block(synth.make_weights(a.small), a.small, dev)
    hid, ctx = (96, 64) if a.small else (128, 128)
    net = torch.tanh(torch.randn(B, h, w, hid, device=dev)); inp = torch.relu(torch.randn(B, h, w, ctx, device=dev))
    capi.check(lib.rb_update_set_state(s, capi.ptr(blob), capi.ptr(ws), capi.ptr(net), capi.ptr(inp), B, h, w, capi.stream()))
    c1 = coords.clone()
    mask = torch.empty(B * h * w * 576, device=dev)
    fn = lambda: capi.check(lib.rb_raft_iterate(s, capi.ptr(blob), capi.ptr(ws), capi.ptr(pyr), capi.ptr(c1),
                                                None if a.small else capi.ptr(mask), B, h, w, 4, capi.stream()))
elif a.what == "corr":  # rb_corr_build: 4 GEMMs (volume + pooled levels) + split / pool passes
    C = 128 if a.small else 256
    f1 = torch.randn(B, h, w, C, device=dev); f2 = torch.randn(B, h, w, C, device=dev)
    cwb = capi.size_query(lib.rb_corr_workspace_bytes, B, h, w, C)
    cws = torch.zeros(cwb, dtype=torch.uint8, device=dev)
    fn = lambda: capi.check(lib.rb_corr_build(capi.ptr(f1), capi.ptr(f2), capi.ptr(pyr), B, h, w, C, capi.ptr(cws), cwb, capi.stream()))
elif a.what in ("encoder", "forward"):
    from types import SimpleNamespace
    from networks.RAFT import RAFT
    H, W = h * 8, w * 8
    m = RAFT((H, W, 3), SimpleNamespace(small=a.small), iters=32 if not a.small else 20, batch=B, device=dev).load(synth.make_weights(a.small))
    l, r_ = synth.make_batch(B, H, W)
    ld, rd = torch.from_numpy(l).to(dev), torch.from_numpy(r_).to(dev)
    eng = m.engine()
    eng.forward(ld, rd)
    if a.what == "encoder":
        eng.use_graph = False
        fn = lambda: (eng.encode(), eng._join_cnet())
    else:
        fn = lambda: eng.forward(ld, rd)
        a.n = 3
else:
    raise SystemExit("unknown")
for _ in range(a.reps):
    fn()
torch.cuda.synchronize()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if a.flush else None


def graph(with_fn):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(a.n):
            if flush is not None:
                flush.zero_()
            if with_fn:
                fn()
    return gr


def run(gr):
    ts = []
    gr.replay(); torch.cuda.synchronize()
    for _ in range(5):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); gr.replay(); ev[1].record(); torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    return sorted(ts)[2]


if a.what == "forward":  # the engine replays its own graph
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        if flush is not None:
            flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    print(f"forward: {sorted(ts)[3] * 1e3:.1f} us per call ({'L2 flushed' if a.flush else 'L2 warm'}, B={B})")
else:
    t = run(graph(True)) - (run(graph(False)) if a.flush else 0.0)
    print(f"{a.what}: {t / a.n * 1e3:.2f} us per call ({'L2 flushed' if a.flush else 'L2 warm'}, graph of {a.n}, B={B})")
