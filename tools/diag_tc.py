"""GPU bring-up diagnostics (run under gpurun): exercises each back end on a few conv shapes and prints
error statistics instead of asserting, so one GPU call tells which stage is wrong."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
import numpy as np
import torch

from oracle import raft_oracle as O
from raft_b200 import capi

dev = torch.device("cuda:0")
lib = capi.lib


def conv_case(mode, B, h, w, cin, cout, kh, kw, ones=False):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, h, w, cin, generator=g)
    W = torch.randn(kh, kw, cin, cout, generator=g) * (2.0 / (kh * kw * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    if ones:
        x = torch.ones_like(x); W = torch.ones_like(W) / (kh * kw * cin); b = torch.zeros_like(b)
    ref = O.conv2d(x.double(), W.double(), b.double(), 1, None)
    y = torch.full((B, h, w, cout), float("nan"), device=dev)
    wsb = capi.size_query(lib.rb_conv2d_workspace_bytes, B, h, w, cin, cout, kh, kw)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    Wn, bn = np.ascontiguousarray(W.numpy()), np.ascontiguousarray(b.numpy())
    capi.check(lib.rb_set_math_mode(mode))
    capi.check(lib.rb_conv2d(capi.ptr(x.to(dev)), Wn.ctypes.data, bn.ctypes.data, capi.ptr(y), B, h, w, cin, cout, kh, kw, 0,
                             capi.ptr(ws), wsb, capi.stream()))
    torch.cuda.synchronize()
    yc = y.cpu().double()
    nan = int((~torch.isfinite(yc)).sum())
    d = (torch.nan_to_num(yc) - ref).abs()
    print(f"  mode={mode} {B}x{h}x{w} cin={cin} cout={cout} k={kh}x{kw} ones={ones}: nan={nan} maxerr={d.max():.3e} "
          f"meanerr={d.mean():.3e} refmax={ref.abs().max():.3f}")
    if d.max() > 1e-3 and not nan:
        idx = torch.nonzero(d > 1e-3)
        print("   first bad idx:", idx[:5].tolist(), "count", len(idx), "of", d.numel())
        # per-pixel / per-channel pattern
        print("   bad per channel (first 16):", (d > 1e-3).sum(dim=(0, 1, 2))[:16].tolist())
        print("   bad per row y:", (d > 1e-3).sum(dim=(0, 2, 3)).tolist()[:16])
        print("   bad per col x (first 16):", (d > 1e-3).sum(dim=(0, 1, 3)).tolist()[:16])
        print("   sample out/ref:", yc.flatten()[:6].tolist(), ref.flatten()[:6].tolist())


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    for mode in (capi.RB_MATH_SIMT, capi.RB_MATH_TC):
        for case in [(1, 8, 16, 64, 16, 1, 1), (1, 8, 16, 64, 64, 1, 1), (1, 8, 16, 128, 128, 1, 1), (1, 8, 16, 64, 64, 3, 3),
                     (1, 16, 32, 384, 256, 1, 5), (2, 13, 27, 242, 192, 3, 3)]:
            for ones in (True, False):
                try:
                    conv_case(mode, *case, ones=ones)
                except Exception:
                    traceback.print_exc()
                    break
