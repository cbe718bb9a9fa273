"""Host-side weight packing (no GPU): the blobs rb_update_weights_pack / rb_encoder_weights_pack upload, produced by the
host-only C-ABI forms and checked against the reference's semantics on the CPU oracle:

  * every conv as split fp16 planes [cout_pad][kh*kw][cin_pad]: hi + 2^-11 lo reproduces the fp32 weight to ~2^-22, the
    convz | convr pairs are concatenated along cout, padding is zero;
  * batch norm folded into W / b (inference statistics, eps 1e-5) == conv followed by the norm (model_utils.py:6-16);
  * the 7x7 stride-2 stem as a 4x1 conv over the zero-padded space-to-depth view (csrc/encoder.cu) and the update block's
    7x7x2 convf1 as a 7x1 conv over the 8-pixel window view (csrc/update.cu): a numpy walk over those views with the
    PACKED weights equals the oracle's plain convolution (TF 'SAME' offsets, even and odd sizes)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))

from oracle import raft_oracle as O  # noqa: E402
from raft_b200 import capi, synth  # noqa: E402

lib = capi.lib
LO = 2.0 ** -11


def _planes(blob, hi, lo, bias, cout_pad, taps, cin_pad):
    n = cout_pad * taps * cin_pad
    h = np.frombuffer(blob, dtype=np.float16, count=n, offset=hi).astype(np.float64)
    l = np.frombuffer(blob, dtype=np.float16, count=n, offset=lo).astype(np.float64)
    b = np.frombuffer(blob, dtype=np.float32, count=cout_pad, offset=bias).astype(np.float64)
    return (h + LO * l).reshape(cout_pad, taps, cin_pad), b


def _update_blob(small):
    p = synth.make_weights(small)
    s = int(small)
    n = lib.rb_update_num_convs(s)
    Ws, bs, keep, names = (C.c_void_p * n)(), (C.c_void_p * n)(), [], []
    for i in range(n):
        name = lib.rb_update_conv_name(s, i).decode()
        W = np.ascontiguousarray(p[name + "/W"], dtype=np.float32)
        b = np.ascontiguousarray(p[name + "/b"], dtype=np.float32)
        keep += [W, b]
        names.append(name)
        Ws[i], bs[i] = W.ctypes.data, b.ctypes.data
    nbytes = capi.size_query(lib.rb_update_weights_bytes, s)
    blob = np.zeros(nbytes, dtype=np.uint8)
    capi.check(lib.rb_update_weights_pack_host(s, Ws, bs, blob.ctypes.data, nbytes))
    return p, blob, names


def _update_conv(small, blob, cid):
    hi, lo, bias = C.c_size_t(), C.c_size_t(), C.c_size_t()
    kh, kw, cin_pad, cout, cout_pad = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    capi.check(lib.rb_update_packed_conv(int(small), cid, C.byref(hi), C.byref(lo), C.byref(bias), C.byref(kh), C.byref(kw),
                                         C.byref(cin_pad), C.byref(cout), C.byref(cout_pad)))
    if cout.value == 0:
        return None
    W, b = _planes(blob, hi.value, lo.value, bias.value, cout_pad.value, kh.value * kw.value, cin_pad.value)
    return W, b, kh.value, kw.value, cout.value


# packed id -> reference conv(s), in the order documented in include/raft_b200.h
THINGS = {0: ["encoder/convc1"], 1: ["encoder/convc2"], 2: ["encoder/convf2"], 3: ["encoder/conv"],
          4: ["gru/convz1", "gru/convr1"], 5: ["gru/convq1"], 6: ["gru/convz2", "gru/convr2"], 7: ["gru/convq2"],
          8: ["flow_head/conv1"], 9: ["flow_head/conv2"], 10: ["mask/0"], 11: ["mask/2"]}
SMALL = {0: ["encoder/convc1"], 2: ["encoder/convf2"], 3: ["encoder/conv"], 4: ["gru/convz", "gru/convr"], 5: ["gru/convq"],
         8: ["flow_head/conv1"], 9: ["flow_head/conv2"]}


@pytest.mark.parametrize("small", [False, True])
def test_update_block_packing(small):
    p, blob, names = _update_blob(small)
    table = SMALL if small else THINGS
    seen = 0
    for cid in range(12):
        got = _update_conv(small, blob, cid)
        if cid not in table:
            assert got is None, f"packed conv {cid} should not exist for small={small}"
            continue
        W, b, kh, kw, cout = got
        srcs = ["update_block/" + n for n in table[cid]]
        for n in srcs:
            assert n in names, n
        Wref = np.concatenate([p[n + "/W"] for n in srcs], axis=3).astype(np.float64)  # HWIO, concatenated along cout
        bref = np.concatenate([p[n + "/b"] for n in srcs]).astype(np.float64)
        assert (kh, kw) == Wref.shape[:2] and cout == Wref.shape[3]
        cin = Wref.shape[2]
        want = Wref.reshape(kh * kw, cin, cout).transpose(2, 0, 1)  # [cout][tap][cin]
        assert np.abs(W[:cout, :, :cin] - want).max() <= 2.0 ** -21 * max(np.abs(want).max(), 1e-3)
        assert np.all(W[cout:] == 0) and np.all(W[:, :, cin:] == 0), "padding must be zero"
        assert np.array_equal(b[:cout], bref.astype(np.float32).astype(np.float64)) and np.all(b[cout:] == 0)
        seen += 1
    assert seen == len(table)


@pytest.mark.parametrize("small", [False, True])
@pytest.mark.parametrize("h,w", [(9, 16), (7, 13)])
def test_convf1_window_view_equals_7x7_conv(small, h, w):
    """csrc/update.cu flow_prep_kernel + the 7x1 conv over the 8-pixel window view, walked in numpy with the PACKED weights."""
    p, blob, _ = _update_blob(small)
    W, b, kh, kw, cout = _update_conv(small, blob, 100)
    assert (kh, kw) == (7, 1) and W.shape[1:] == (7, 64)
    g = torch.Generator().manual_seed(5)
    flow = torch.randn(1, h, w, 2, generator=g, dtype=torch.float64) * 3
    ref = O.conv2d(flow, torch.from_numpy(p["update_block/encoder/convf1/W"]).double(),
                   torch.from_numpy(p["update_block/encoder/convf1/b"]).double(), 1, torch.relu)[0].numpy()
    fl = np.zeros((h, w + 8, 8))  # 3 zero pixels left, 5 right, channels 0,1 = flow
    fl[:, 3:3 + w, :2] = flow[0].numpy()
    out = np.zeros((h, w, cout))
    for y in range(h):
        for x in range(w):
            acc = b[:cout].copy()
            for ky in range(7):
                yy = y + ky - 3  # pad_y = 3, rows outside the image are TMA zero fill
                if 0 <= yy < h:
                    acc += W[:cout, ky, :] @ fl[yy, x:x + 8, :].reshape(64)  # view pixel x = physical pixels x .. x+7
            out[y, x] = np.maximum(acc, 0.0)
    assert np.abs(out - ref).max() < 1e-5 * max(np.abs(ref).max(), 1.0)


def _encoder_blob(small, name, norm):
    p = synth.make_weights(small)
    s, out_dim = int(small), (128 if name == "fnet" and small else 256 if name == "fnet" else 160 if small else 256)
    n = lib.rb_encoder_num_convs(s)
    Ws, bs, bns, keep, convs = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)(), [], []
    for i in range(n):
        cname = f"{name}/{lib.rb_encoder_conv_name(s, i).decode()}"
        nname = lib.rb_encoder_norm_name(s, i).decode()
        W = np.ascontiguousarray(p[cname + "/W"], dtype=np.float32)
        b = np.ascontiguousarray(p[cname + "/b"], dtype=np.float32)
        keep += [W, b]
        Ws[i], bs[i] = W.ctypes.data, b.ctypes.data
        bn = None
        if norm == 2 and nname:
            sc = f"{name}/{nname}"
            bn = np.ascontiguousarray(np.concatenate([p[sc + "/gamma"], p[sc + "/beta"], p[sc + "/mean/EMA"],
                                                      p[sc + "/variance/EMA"]]), dtype=np.float32)
            keep.append(bn)
            bns[i] = bn.ctypes.data
        convs.append((cname, W, b, bn))
    nbytes = capi.size_query(lib.rb_encoder_weights_bytes, s, out_dim)
    blob = np.zeros(nbytes, dtype=np.uint8)
    capi.check(lib.rb_encoder_weights_pack_host(s, norm, out_dim, Ws, bs, bns, blob.ctypes.data, nbytes))
    return blob, convs, out_dim


def _encoder_conv(small, out_dim, blob, i):
    hi, lo, bias = C.c_size_t(), C.c_size_t(), C.c_size_t()
    kh, kw, cin_pad, cout_pad = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    capi.check(lib.rb_encoder_packed_conv(int(small), out_dim, i, C.byref(hi), C.byref(lo), C.byref(bias), C.byref(kh),
                                          C.byref(kw), C.byref(cin_pad), C.byref(cout_pad)))
    W, b = _planes(blob, hi.value, lo.value, bias.value, cout_pad.value, kh.value * kw.value, cin_pad.value)
    return W, b, kh.value, kw.value


@pytest.mark.parametrize("small,name,norm", [(False, "cnet", 2), (False, "fnet", 1), (True, "cnet", 0)])
def test_encoder_packing_and_batch_norm_fold(small, name, norm):
    blob, convs, out_dim = _encoder_blob(small, name, norm)
    for i, (cname, Wsrc, bsrc, bn) in enumerate(convs[1:], start=1):  # the stem has its own test
        W, b, kh, kw = _encoder_conv(small, out_dim, blob, i)
        k, _, cin, cout = Wsrc.shape
        assert (kh, kw) == (k, k), cname
        scale, shift = np.ones(cout), np.zeros(cout)
        if bn is not None:  # y = (conv + b - mean) / sqrt(var + eps) * gamma + beta
            gamma, beta, mean, var = bn.astype(np.float64).reshape(4, cout)
            scale = gamma / np.sqrt(var + 1e-5)
            shift = beta - mean * scale
        want = (Wsrc.astype(np.float64) * scale).reshape(k * k, cin, cout).transpose(2, 0, 1)
        assert np.abs(W[:cout, :, :cin] - want).max() <= 2.0 ** -20 * max(np.abs(want).max(), 1e-3), cname
        assert np.all(W[cout:] == 0) and np.all(W[:, :, cin:] == 0), cname
        assert np.abs(b[:cout] - (bsrc.astype(np.float64) * scale + shift)).max() <= 1e-6 * max(np.abs(shift).max(), 1.0), cname


@pytest.mark.parametrize("small,name,norm", [(False, "cnet", 2), (True, "fnet", 1)])
@pytest.mark.parametrize("H,W_", [(16, 24), (15, 21)])
def test_stem_space_to_depth_view_equals_7x7_stride2_conv(small, name, norm, H, W_):
    """csrc/encoder.cu enc_stem_s2d_kernel + the 4x1 conv over the overlapping window view, walked in numpy with the PACKED
    (batch-norm-folded) weights, against the oracle's 7x7 stride-2 'SAME' conv followed by the norm."""
    blob, convs, out_dim = _encoder_blob(small, name, norm)
    cname, Wsrc, bsrc, bn = convs[0]
    W, b, kh, kw = _encoder_conv(small, out_dim, blob, 0)
    assert (kh, kw) == (4, 1) and W.shape[2] == 64
    cout = Wsrc.shape[3]
    g = torch.Generator().manual_seed(9)
    img = torch.rand(1, H, W_, 3, generator=g, dtype=torch.float64) * 2 - 1  # already 2x-1
    ref = O.conv2d(img, torch.from_numpy(Wsrc).double(), torch.from_numpy(bsrc).double(), 2, None)[0].numpy()
    if bn is not None:
        gamma, beta, mean, var = bn.astype(np.float64).reshape(4, cout)
        ref = (ref - mean) / np.sqrt(var + 1e-5) * gamma + beta
    oh, ow = (H + 1) // 2, (W_ + 1) // 2
    pt = max((oh - 1) * 2 + 7 - H, 0) // 2  # TF SAME: pad before = total / 2
    pl = max((ow - 1) * 2 + 7 - W_, 0) // 2
    Hp, Wp = oh + 3, ow + 3
    cells = np.zeros((Hp, Wp, 16))  # cell (Y, X) = padded rows 2Y, 2Y+1 x cols 2X, 2X+1, channel (dy*2+dx)*3 + c
    x = img[0].numpy()
    for Y in range(Hp):
        for X in range(Wp):
            for dy in range(2):
                for dx in range(2):
                    r, q = 2 * Y + dy - pt, 2 * X + dx - pl
                    if 0 <= r < H and 0 <= q < W_:
                        cells[Y, X, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = x[r, q]
    out = np.zeros((oh, ow, cout))
    for oy in range(oh):
        for ox in range(ow):
            acc = b[:cout].copy()
            for ky in range(4):  # pad 0: cell rows oy .. oy+3; the window = cells ox .. ox+3 = 64 contiguous channels
                acc += W[:cout, ky, :] @ cells[oy + ky, ox:ox + 4, :].reshape(64)
            out[oy, ox] = acc
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() < 2e-5 * max(np.abs(ref).max(), 1.0)
