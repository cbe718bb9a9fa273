"""CPU oracle for the RAFT recurrent-inference hot path (TEST INFRASTRUCTURE ONLY).

This file is a plain torch-CPU restatement (fp32 by default, fp64 on request) of the
algorithm in gonglixue/RAFT-tf.  It exists so that the CUDA path can be checked against
something; it is NOT part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

PARITY STATUS: the reference (TensorFlow 1.15 + tensorpack) cannot be executed in this
environment and ships no tests or golden vectors, so this oracle is **not pinned against
real TensorFlow output** ("parity unpinned" in that strict sense).  What pins it instead:
``oracle/ref_shim`` executes the reference's OWN source files (``networks/utils.py``,
``networks/model_utils.py``, ``networks/RAFT.py``) on a numpy stand-in for the handful of
``tf.*`` / tensorpack calls they make, and ``tests/golden/*.npz`` holds the vectors that
run produced; ``tests/test_oracle_vs_refshim.py`` compares this file against them.

Every function cites the reference lines it follows (paths relative to the reference
root).  All tensors are NHWC like the reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# Error-budget experiments only (tests/test_precision_budget.py): when set,
# conv2d / the correlation matmul round their operands the way a tensor-core path would.
# None = plain fp32/fp64 arithmetic (the oracle proper).
EMULATE: Optional[str] = None


def _split_terms(a: Tensor, mode: str):
    """Return [(a_part, scale)] such that a ~= sum(part*scale) under the emulated format."""
    if mode == "tf32":
        ai = a.float().view(torch.int32)
        ai = (ai + 0x1000) & ~0x1FFF  # round-to-nearest (ties away) to 10 mantissa bits
        return [(ai.view(torch.float32).to(a.dtype), 1.0)]
    if mode == "bf16":
        return [(a.bfloat16().to(a.dtype), 1.0)]
    if mode == "fp16":
        return [(a.half().to(a.dtype), 1.0)]
    if mode == "bf16x3":
        hi = a.bfloat16().to(a.dtype)
        lo = (a - hi).bfloat16().to(a.dtype)
        return [(hi, 1.0), (lo, 1.0)]
    if mode == "fp16x3":  # hi + 2^-11 * lo', lo' kept in fp16 at 2^11 scale
        hi = a.half().to(a.dtype)
        lo = ((a - hi) * 2048.0).half().to(a.dtype)
        return [(hi, 1.0), (lo, 1.0 / 2048.0)]
    raise ValueError(mode)


def _emulated_bilinear(op, a: Tensor, b: Tensor) -> Tensor:
    """op(a,b) bilinear; drop the lo*lo term for the 3-pass modes."""
    A, B = _split_terms(a, EMULATE), _split_terms(b, EMULATE)
    out = op(A[0][0], B[0][0])
    if len(A) > 1:
        cross = op(A[0][0], B[1][0]) + op(A[1][0], B[0][0])
        out = out + cross * A[1][1]
    return out


# --------------------------------------------------------------------------------------
# third-party layer semantics the reference relies on (tensorpack / TF documented defaults)
# --------------------------------------------------------------------------------------
def _same_pad(n: int, k: int, s: int) -> Tuple[int, int]:
    """TF 'SAME' padding: out=ceil(n/s); total=max((out-1)*s+k-n,0); before=total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d(x: Tensor, W: Tensor, b: Optional[Tensor], stride: int = 1, activation=None) -> Tensor:
    """tensorpack ``Conv2D(name, x, filters, kernel_size, strides, padding='same')`` +bias
    (+activation).  x NHWC, W HWIO ``[kh,kw,cin,cout]``.  (model_utils.py:21,25,... A14)"""
    kh, kw = int(W.shape[0]), int(W.shape[1])
    pt, pb = _same_pad(x.shape[1], kh, stride)
    pl, pr = _same_pad(x.shape[2], kw, stride)
    xn = x.permute(0, 3, 1, 2)
    xn = F.pad(xn, (pl, pr, pt, pb))
    Wn = W.permute(3, 2, 0, 1).contiguous()
    if EMULATE is None:
        y = F.conv2d(xn, Wn, b, stride=stride)
    else:
        y = _emulated_bilinear(lambda u, v: F.conv2d(u, v, None, stride=stride), xn, Wn)
        if b is not None:
            y = y + b[None, :, None, None]
    y = y.permute(0, 2, 3, 1).contiguous()
    if activation is not None:
        y = activation(y)
    return y


def avg_pool_2x2_valid(x: Tensor) -> Tensor:
    """tensorpack ``AvgPooling(pool_size=2, strides=2)`` default padding 'valid'
    (floor on odd dims).  x NHWC.  (model_utils.py:218)"""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


def instance_norm(x: Tensor, eps: float = 1e-5) -> Tensor:
    """tensorpack ``InstanceNorm(center=False, scale=False)``: per-sample, per-channel
    moments over H,W (biased variance), no affine.  (model_utils.py:13)"""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def batch_norm_inference(x: Tensor, p: Dict[str, Tensor], scope: str, eps: float = 1e-5) -> Tensor:
    """tensorpack ``BatchNorm`` at inference: EMA statistics.  (model_utils.py:11)"""
    g, bt = p[scope + "/gamma"], p[scope + "/beta"]
    mu, var = p[scope + "/mean/EMA"], p[scope + "/variance/EMA"]
    return (x - mu) / torch.sqrt(var + eps) * g + bt


# --------------------------------------------------------------------------------------
# networks/model_utils.py:6-105 -- encoders
# --------------------------------------------------------------------------------------
def _norm(x: Tensor, p, scope: str, norm_fn: str) -> Tensor:
    """model_utils.py:6-17 (the 'group' branch is never selected: RAFT.py:64-74)."""
    if norm_fn == "batch":
        return batch_norm_inference(x, p, scope)
    if norm_fn == "instance":
        return instance_norm(x)
    if norm_fn == "none":
        return x
    raise ValueError(norm_fn)


def _conv(x, p, scope, stride=1, activation=None):
    return conv2d(x, p[scope + "/W"], p[scope + "/b"], stride, activation)


def residual_block(x, p, scope, norm_fn, stride):
    """model_utils.py:19-35."""
    res = x
    y = _conv(x, p, scope + "/conv1", stride)
    y = torch.relu(_norm(y, p, scope + "/norm1", norm_fn))
    y = _conv(y, p, scope + "/conv2", 1)
    y = torch.relu(_norm(y, p, scope + "/norm2", norm_fn))
    if stride != 1:
        res = _conv(res, p, scope + "/downsample.0", stride)
        res = _norm(res, p, scope + "/downsample.1", norm_fn)
    return torch.relu(res + y)


def bottleneck_block(x, p, scope, norm_fn, stride):
    """model_utils.py:37-57."""
    res = x
    y = torch.relu(_norm(_conv(x, p, scope + "/conv1", 1), p, scope + "/norm1", norm_fn))
    y = torch.relu(_norm(_conv(y, p, scope + "/conv2", stride), p, scope + "/norm2", norm_fn))
    y = torch.relu(_norm(_conv(y, p, scope + "/conv3", 1), p, scope + "/norm3", norm_fn))
    if stride != 1:
        res = _conv(res, p, scope + "/downsample.0", stride)
        res = _norm(res, p, scope + "/downsample.1", norm_fn)
    return torch.relu(res + y)


def basic_encoder(x, p, name, norm_fn):
    """model_utils.py:61-82 (dropout is dead at inference)."""
    y = torch.relu(_norm(_conv(x, p, name + "/conv1", 2), p, name + "/norm1", norm_fn))
    for lname, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        y = residual_block(y, p, f"{name}/{lname}/0", norm_fn, stride)
        y = residual_block(y, p, f"{name}/{lname}/1", norm_fn, 1)
    return _conv(y, p, name + "/conv2", 1)


def small_encoder(x, p, name, norm_fn):
    """model_utils.py:84-105."""
    y = torch.relu(_norm(_conv(x, p, name + "/conv1", 2), p, name + "/norm1", norm_fn))
    for lname, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        y = bottleneck_block(y, p, f"{name}/{lname}/0", norm_fn, stride)
        y = bottleneck_block(y, p, f"{name}/{lname}/1", norm_fn, 1)
    return _conv(y, p, name + "/conv2", 1)


# --------------------------------------------------------------------------------------
# networks/utils.py -- grid, sampler, upflow8
# --------------------------------------------------------------------------------------
def coords_grid(batch: int, ht: int, wd: int, dtype=torch.float32) -> Tensor:
    """utils.py:4-11: [b,h,w,2], channel 0 = x (column), 1 = y (row)."""
    xs, ys = torch.meshgrid(torch.arange(wd), torch.arange(ht), indexing="xy")
    g = torch.stack([xs, ys], dim=-1).to(dtype)
    return g[None].repeat(batch, 1, 1, 1)


def tf_grid_sample(img: Tensor, coords: Tensor) -> Tensor:
    """utils.py:39-99.  img [n,H,W,1], coords [n,a,b,2] (x,y) in pixels.
    tf.cast(float->int32) truncates toward zero (:54-57); indices are clamped (:60-63);
    the weights use the CLAMPED x1,y1 (:84-89); out = wa*Ia+wb*Ib+wc*Ic+wd*Id (:98)."""
    n, H, W, _ = img.shape
    x = coords[..., 0]
    y = coords[..., 1]
    x0 = torch.trunc(x).to(torch.int64)
    y0 = torch.trunc(y).to(torch.int64)
    x1 = x0 + 1
    y1 = y0 + 1
    x0 = x0.clamp(0, W - 1)
    x1 = x1.clamp(0, W - 1)
    y0 = y0.clamp(0, H - 1)
    y1 = y1.clamp(0, H - 1)
    flat = img.reshape(n, H * W)

    def gather(xi, yi):  # get_pixel_value, utils.py:13-37
        idx = (yi * W + xi).reshape(n, -1)
        return torch.gather(flat, 1, idx).reshape(x.shape)

    Ia = gather(x0, y0)
    Ib = gather(x0, y1)
    Ic = gather(x1, y0)
    Id = gather(x1, y1)
    qx = x1.to(img.dtype) - x
    qy = y1.to(img.dtype) - y
    wa = qx * qy
    wb = qx * (1.0 - qy)
    wc = (1.0 - qx) * qy
    wd = (1.0 - qx) * (1.0 - qy)
    out = ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id  # tf.add_n, left to right
    return out[..., None]


def bilinear_sampler(img, coords):
    """utils.py:101-103."""
    return tf_grid_sample(img, coords)


def resize_bilinear_align_corners(x: Tensor, oh: int, ow: int) -> Tensor:
    """tf.image.resize_bilinear(align_corners=True) (TF resize_bilinear_op.cc semantics):
    scale=(in-1)/(out-1); in=out_idx*scale; lower=floor; upper=min(lower+1,in-1);
    top=tl+(tr-tl)*xl; bottom=bl+(br-bl)*xl; out=top+(bottom-top)*yl.  x NHWC."""
    b, h, w, c = x.shape
    sy = (h - 1) / (oh - 1) if oh > 1 else 0.0
    sx = (w - 1) / (ow - 1) if ow > 1 else 0.0
    fy = torch.arange(oh, dtype=torch.float32) * np.float32(sy)
    fx = torch.arange(ow, dtype=torch.float32) * np.float32(sx)
    y0 = torch.floor(fy).long()
    x0 = torch.floor(fx).long()
    y1 = (y0 + 1).clamp(max=h - 1)
    x1 = (x0 + 1).clamp(max=w - 1)
    ly = (fy - y0.float()).to(x.dtype)[None, :, None, None]
    lx = (fx - x0.float()).to(x.dtype)[None, None, :, None]
    tl = x[:, y0][:, :, x0]
    tr = x[:, y0][:, :, x1]
    bl = x[:, y1][:, :, x0]
    br = x[:, y1][:, :, x1]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    return top + (bot - top) * ly


def upflow8(flow: Tensor) -> Tensor:
    """utils.py:105-111: bilinear x8 resize, align_corners=True, and -- unlike upstream
    RAFT -- NO multiplication by 8 (reference quirk; replicated)."""
    b, h, w, _ = flow.shape
    return resize_bilinear_align_corners(flow, 8 * h, 8 * w)


# --------------------------------------------------------------------------------------
# networks/model_utils.py:199-249 -- correlation pyramid and lookup
# --------------------------------------------------------------------------------------
def get_corr_pyramid(fmap1: Tensor, fmap2: Tensor, num_levels: int = 4) -> List[Tensor]:
    """model_utils.py:199-221 (GetCorrPyramid).  Returns levels [B*h*w, h_i, w_i, 1]."""
    b, h, w, c = fmap1.shape
    f1 = fmap1.reshape(b, h * w, c)
    f2 = fmap2.reshape(b, h * w, c)
    if EMULATE is None:
        corr = torch.matmul(f1, f2.transpose(1, 2))  # :209-210
    else:
        corr = _emulated_bilinear(lambda u, v: torch.matmul(u, v.transpose(1, 2)), f1, f2)
    corr = corr / math.sqrt(float(c)) if fmap1.dtype == torch.float64 else corr / torch.sqrt(
        torch.tensor(float(c), dtype=fmap1.dtype))  # :213, divide AFTER the matmul
    corr = corr.reshape(b * h * w, h, w, 1)  # :215
    pyr = [corr]
    for _ in range(num_levels - 1):
        corr = avg_pool_2x2_valid(corr)  # :217-219
        pyr.append(corr)
    return pyr


def sample_corr(pyramid: List[Tensor], coords: Tensor, num_levels: int = 4, radius: int = 4) -> Tensor:
    """model_utils.py:224-249 (SampleCorr).  coords [b,h,w,2] -> [b,h,w,L*(2r+1)^2].
    delta[i,j] = (x_off=i-r, y_off=j-r): the FIRST window axis walks x (:235-237)."""
    b, h, w, _ = coords.shape
    d = torch.linspace(-float(radius), float(radius), 2 * radius + 1, dtype=coords.dtype)
    # tf.meshgrid(dy,dx) is 'xy' indexing: out0[i,j]=dy[j], out1[i,j]=dx[i]; [::-1] then
    # stack -> delta[i,j] = (dx[i], dy[j])
    delta = torch.stack([d[:, None].expand(-1, 2 * radius + 1), d[None, :].expand(2 * radius + 1, -1)], dim=-1)
    out = []
    for i in range(num_levels):
        centroid = coords.reshape(b * h * w, 1, 1, 2) / (2 ** i)  # :239
        coords_lvl = centroid + delta[None]  # :241
        c = bilinear_sampler(pyramid[i], coords_lvl)  # :244
        out.append(c.reshape(b, h, w, -1))  # :245
    return torch.cat(out, dim=-1)  # :248


# --------------------------------------------------------------------------------------
# networks/model_utils.py:110-194 -- update blocks
# --------------------------------------------------------------------------------------
def basic_motion_encoder(flow, corr, p, name):
    """model_utils.py:110-119."""
    cor = _conv(corr, p, name + "/convc1", 1, torch.relu)
    cor = _conv(cor, p, name + "/convc2", 1, torch.relu)
    flo = _conv(flow, p, name + "/convf1", 1, torch.relu)
    flo = _conv(flo, p, name + "/convf2", 1, torch.relu)
    out = _conv(torch.cat([cor, flo], -1), p, name + "/conv", 1, torch.relu)
    return torch.cat([out, flow], -1)


def small_motion_encoder(flow, corr, p, name):
    """model_utils.py:121-129."""
    cor = _conv(corr, p, name + "/convc1", 1, torch.relu)
    flo = _conv(flow, p, name + "/convf1", 1, torch.relu)
    flo = _conv(flo, p, name + "/convf2", 1, torch.relu)
    out = _conv(torch.cat([cor, flo], -1), p, name + "/conv", 1, torch.relu)
    return torch.cat([out, flow], -1)


def flow_head(x, p, name):
    """model_utils.py:131-135."""
    return _conv(_conv(x, p, name + "/conv1", 1, torch.relu), p, name + "/conv2", 1)


def sep_conv_gru(h, x, p, name):
    """model_utils.py:138-156: (1,5) pass then (5,1) pass."""
    for s in ("1", "2"):
        hx = torch.cat([h, x], -1)
        z = _conv(hx, p, f"{name}/convz{s}", 1, torch.sigmoid)
        r = _conv(hx, p, f"{name}/convr{s}", 1, torch.sigmoid)
        q = _conv(torch.cat([r * h, x], -1), p, f"{name}/convq{s}", 1, torch.tanh)
        h = (1 - z) * h + z * q
    return h


def conv_gru(h, x, p, name):
    """model_utils.py:158-169."""
    hx = torch.cat([h, x], -1)
    z = _conv(hx, p, name + "/convz", 1, torch.sigmoid)
    r = _conv(hx, p, name + "/convr", 1, torch.sigmoid)
    q = _conv(torch.cat([r * h, x], -1), p, name + "/convq", 1, torch.tanh)
    return (1 - z) * h + z * q


def basic_update_block(net, inp, corr, flow, p, name="update_block", with_mask=True):
    """model_utils.py:172-185."""
    motion = basic_motion_encoder(flow, corr, p, name + "/encoder")
    x = torch.cat([inp, motion], -1)
    net = sep_conv_gru(net, x, p, name + "/gru")
    delta = flow_head(net, p, name + "/flow_head")
    mask = None
    if with_mask:
        m = _conv(net, p, name + "/mask/0", 1, torch.relu)
        m = _conv(m, p, name + "/mask/2", 1)
        mask = 0.25 * m
    return net, mask, delta


def small_update_block(net, inp, corr, flow, p, name="update_block"):
    """model_utils.py:187-194."""
    motion = small_motion_encoder(flow, corr, p, name + "/encoder")
    x = torch.cat([inp, motion], -1)
    net = conv_gru(net, x, p, name + "/gru")
    delta = flow_head(net, p, name + "/flow_head")
    return net, None, delta


# --------------------------------------------------------------------------------------
# networks/RAFT.py
# --------------------------------------------------------------------------------------
def upsample_flow(flow: Tensor, mask: Tensor) -> Tensor:
    """RAFT.py:119-134: convex 8x upsampling.  mask channel = k*64 + sy*8 + sx, k=ky*3+kx;
    softmax over k; 3x3 zero-padded patches of 8*flow (depth order ky,kx,c)."""
    b, h, w, _ = flow.shape
    m = mask.reshape(b, h, w, 9, 1, 8, 8)
    m = torch.softmax(m, dim=3)
    f8 = (8 * flow).permute(0, 3, 1, 2)  # NCHW
    fp = F.pad(f8, (1, 1, 1, 1))
    patches = []
    for ky in range(3):
        for kx in range(3):
            patches.append(fp[:, :, ky:ky + h, kx:kx + w].permute(0, 2, 3, 1))  # [b,h,w,2]
    up = torch.stack(patches, dim=3).reshape(b, h, w, 9, 2, 1, 1)
    up = (up * m).sum(dim=3)  # [b,h,w,2,8,8]
    up = up.permute(0, 1, 4, 2, 5, 3)  # [b,h,8,w,8,2]
    return up.reshape(b, h * 8, w * 8, 2)


class RAFTOracle:
    """RAFT.py:12-141 restated.  ``params`` maps reference variable names (npz keys) to
    arrays.  ``forward`` takes [B,H,W,3] images in [0,1] (BGR; H,W multiples of 8)."""

    def __init__(self, params: Dict[str, np.ndarray], small: bool = False, iters: int = 20,
                 dtype=torch.float32):
        self.small = small
        self.iters = iters  # RAFT.py:33
        self.dtype = dtype
        self.hidden_dim, self.context_dim, self.corr_radius = (96, 64, 3) if small else (128, 128, 4)
        self.p = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in params.items()}

    # RAFT.py:62-76
    def feature_extractor(self, img):
        if self.small:
            return small_encoder(img, self.p, "fnet", "instance")
        return basic_encoder(img, self.p, "fnet", "instance")

    def context_net(self, img):
        if self.small:
            return small_encoder(img, self.p, "cnet", "none")
        return basic_encoder(img, self.p, "cnet", "batch")

    def prepare(self, left: Tensor, right: Tensor):
        """RAFT.py:53-59,79-89: everything before the loop."""
        left = 2.0 * left.to(self.dtype) - 1.0
        right = 2.0 * right.to(self.dtype) - 1.0
        fmap1 = self.feature_extractor(left)
        fmap2 = self.feature_extractor(right)
        pyramid = get_corr_pyramid(fmap1, fmap2)
        cnet = self.context_net(left)
        net, inp = torch.split(cnet, [self.hidden_dim, self.context_dim], dim=-1)
        net = torch.tanh(net)
        inp = torch.relu(inp)
        b, H, W, _ = left.shape
        coords0 = coords_grid(b, H // 8, W // 8, self.dtype)
        coords1 = coords_grid(b, H // 8, W // 8, self.dtype)
        return dict(fmap1=fmap1, fmap2=fmap2, pyramid=pyramid, net=net, inp=inp,
                    coords0=coords0, coords1=coords1)

    def iterate(self, st, iters=None, trace=None):
        """RAFT.py:91-102."""
        net, inp, coords0, coords1 = st["net"], st["inp"], st["coords0"], st["coords1"]
        up_mask = None
        iters = self.iters if iters is None else iters
        for it in range(iters):
            corr = sample_corr(st["pyramid"], coords1, radius=self.corr_radius)
            flow = coords1 - coords0
            if self.small:
                net, up_mask, delta = small_update_block(net, inp, corr, flow, self.p)
            else:
                # the mask head only matters in the last iteration (fetch pruning)
                net, up_mask, delta = basic_update_block(net, inp, corr, flow, self.p,
                                                         with_mask=(it == iters - 1))
            coords1 = coords1 + delta
            if trace is not None:
                trace.append(dict(corr=corr, net=net, delta=delta, coords1=coords1))
        return net, up_mask, coords1

    def forward(self, left: Tensor, right: Tensor, iters=None, return_lowres=False):
        st = self.prepare(left, right)
        net, up_mask, coords1 = self.iterate(st, iters)
        lowres = coords1 - st["coords0"]
        if self.small:
            flow_up = upflow8(lowres)  # RAFT.py:104-105
        else:
            flow_up = upsample_flow(lowres, up_mask)  # RAFT.py:106-107
        if return_lowres:
            return flow_up, lowres
        return flow_up
