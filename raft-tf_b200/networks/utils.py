"""Drop-in mirror of the reference's ``networks/utils.py`` on torch CUDA tensors; every function
runs a raft_b200 CUDA kernel through the C ABI (no TF, no CPU fallback).  NHWC like the reference."""
import torch

from raft_b200 import capi


def coords_grid(batch, ht, wd, device=None):
    """utils.py:4-11 -> [b,h,w,2] with channel 0 = x, 1 = y."""
    device = torch.device(device if device is not None else "cuda")
    out = torch.empty(int(batch), int(ht), int(wd), 2, dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        capi.check(capi.lib.rb_coords_grid(capi.ptr(out), int(batch), int(ht), int(wd), capi.stream()))
    return out


def bilinear_sampler(img, coords):
    """utils.py:101-103 (tf_grid_sample, :39-99): img [n,H,W,1], coords [n,a,b,2] (x,y pixels)
    -> [n,a,b,1].  Truncation toward zero, index clamping, weights from the clamped x1/y1."""
    n, H, W, c = img.shape
    assert c == 1, "the reference only samples the single-channel correlation volume"
    a, b = coords.shape[1], coords.shape[2]
    img, coords = img.contiguous().float(), coords.contiguous().float()
    out = torch.empty(n, a, b, 1, dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        capi.check(capi.lib.rb_bilinear_sample(capi.ptr(img), capi.ptr(coords), capi.ptr(out), n, H, W, a * b,
                                               capi.stream()))
    return out


def upflow8(flow):
    """utils.py:105-111: bilinear x8 (align_corners=True) WITHOUT multiplying by 8 (reference quirk)."""
    b, h, w, _ = flow.shape
    coords1 = (flow + coords_grid(b, h, w, flow.device)).contiguous()
    out = torch.empty(b, 8 * h, 8 * w, 2, dtype=torch.float32, device=flow.device)
    with torch.cuda.device(out.device):
        capi.check(capi.lib.rb_upflow8(capi.ptr(coords1), capi.ptr(out), b, h, w, 1.0, capi.stream()))
    return out
