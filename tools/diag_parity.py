#!/usr/bin/env python
"""Parity diagnosis at a BASELINE config: error of the GPU flow against the fp32 AND the fp64 CPU oracle, plus the fp32 oracle's
own distance from fp64 (how much of a difference is recurrent amplification of rounding noise rather than a defect).
    python tools/diag_parity.py [--small] [--H 540 --W 960 --iters 32] [--simt]     (env knobs select kernel variants)
The oracle results are cached in /tmp so that several variants can be compared in one gpurun call."""
import argparse
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
import numpy as np
import torch
from oracle import raft_oracle as O
from raft_b200 import capi, synth
from networks.RAFT import RAFT

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
ap.add_argument("--H", type=int, default=540)
ap.add_argument("--W", type=int, default=960)
ap.add_argument("--iters", type=int, default=32)
ap.add_argument("--seed0", type=int, default=1003)
ap.add_argument("--simt", action="store_true")
ap.add_argument("--tag", default="")
a = ap.parse_args()
p = synth.make_weights(a.small)
l, r = synth.make_batch(1, a.H, a.W, seed0=a.seed0)
ph, pw = (-a.H) % 8, (-a.W) % 8
pad = lambda x: np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)), mode="edge")
crop = lambda f: f[:, ph // 2:ph // 2 + a.H, pw // 2:pw // 2 + a.W]
cache = f"/tmp/diag_oracle_{int(a.small)}_{a.H}_{a.W}_{a.iters}_{a.seed0}.npz"
if os.path.exists(cache):
    z = np.load(cache); r32, r64 = torch.from_numpy(z["r32"]), torch.from_numpy(z["r64"])
else:
    lt, rt = torch.from_numpy(pad(l)), torch.from_numpy(pad(r))
    r32 = crop(O.RAFTOracle(p, small=a.small, iters=a.iters).forward(lt, rt))
    r64 = crop(O.RAFTOracle(p, small=a.small, iters=a.iters, dtype=torch.float64).forward(lt, rt))
    np.savez(cache, r32=r32.numpy(), r64=r64.numpy())
    print(f"oracle fp32 vs fp64: {(r32.double() - r64).abs().max():.3e}   max|flow| {r64.abs().max():.2f} px")
m = RAFT((a.H, a.W, 3), SimpleNamespace(small=a.small), iters=a.iters, device="cuda:0").load(p)
if a.simt:
    m.engine().math_mode = capi.RB_MATH_SIMT
out = m.forward(l, r).cpu()
e32, e64 = (out - r32).abs(), (out.double() - r64).abs()
bad = (e64.max(-1).values[0] > 1e-3).nonzero()
if len(bad):
    ys, xs = bad[:, 0], bad[:, 1]
    print(f"   {len(bad)} pixels above 1e-3 ({100.0 * len(bad) / (a.H * a.W):.4f} %): rows {int(ys.min())}..{int(ys.max())}, cols {int(xs.min())}..{int(xs.max())}; "
          f"coarse cells {sorted(set((int(y) // 8, int(x) // 8) for y, x in bad.tolist()))[:12]}")
    q = torch.quantile(e64.max(-1).values.flatten()[::7].float(), torch.tensor([0.5, 0.99, 0.9999]))
    print(f"   error quantiles (50 / 99 / 99.99 %): {q[0]:.2e} / {q[1]:.2e} / {q[2]:.2e}")
print(f"{a.tag or 'default':>24}: vs fp32 oracle {e32.max():.3e} (mean {e32.mean():.2e})   vs fp64 oracle {e64.max():.3e} (mean {e64.mean():.2e})")
