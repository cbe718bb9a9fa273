"""F4 end to end on the CPU: a checkpoint in the UPSTREAM (PyTorch RAFT) layout -- real ``torch.nn`` modules with the
published architecture, their genuine ``state_dict()`` keys ('module.' prefix of DataParallel, BatchNorm bookkeeping entries,
the norm of a strided block registered twice as ``norm3`` and ``downsample.1``) saved with ``torch.save`` -- goes through the
converter CLI into the reference's ``.npz`` naming, and the oracle (the restatement of the reference's TF graph) must then
compute what the torch modules compute.  That pins the key mapping, OIHW -> HWIO, the BatchNorm statistics and, as a side
effect, the oracle's block structure against an independent implementation.  Further down: the oracle's correlation pyramid
+ lookup against an upstream-style CorrBlock on torch.grid_sample, convex upsampling against the F.unfold form, upflow8 against
F.interpolate(align_corners=True) -- the TF ops whose semantics the numpy shim (oracle/ref_shim) had to assume.

Sizes: every stride-2 conv sees an ODD extent, where PyTorch's symmetric ``padding = k // 2`` and TensorFlow's ``SAME``
(pad before = total / 2) coincide; on even extents the TF port samples one pixel later than upstream by construction."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))

from oracle import raft_oracle as O  # noqa: E402
from raft_b200 import convert  # noqa: E402
from raft_b200.weights import load_npz  # noqa: E402


def _norm(kind, c):
    return nn.BatchNorm2d(c) if kind == "batch" else nn.InstanceNorm2d(c) if kind == "instance" else nn.Sequential()


class ResidualBlock(nn.Module):
    def __init__(self, cin, planes, norm, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1)
        self.norm1, self.norm2 = _norm(norm, planes), _norm(norm, planes)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(norm, planes)
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = F.relu(self.norm1(self.conv1(x)))
        y = F.relu(self.norm2(self.conv2(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return F.relu(x + y)


class BasicEncoder(nn.Module):
    def __init__(self, out_dim, norm):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        self.norm1 = _norm(norm, 64)
        self.layer1 = nn.Sequential(ResidualBlock(64, 64, norm, 1), ResidualBlock(64, 64, norm, 1))
        self.layer2 = nn.Sequential(ResidualBlock(64, 96, norm, 2), ResidualBlock(96, 96, norm, 1))
        self.layer3 = nn.Sequential(ResidualBlock(96, 128, norm, 2), ResidualBlock(128, 128, norm, 1))
        self.conv2 = nn.Conv2d(128, out_dim, 1)

    def forward(self, x):
        x = F.relu(self.norm1(self.conv1(x)))
        return self.conv2(self.layer3(self.layer2(self.layer1(x))))


class MotionEncoder(nn.Module):
    def __init__(self, cor_planes):
        super().__init__()
        self.convc1 = nn.Conv2d(cor_planes, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def forward(self, flow, corr):
        cor = F.relu(self.convc2(F.relu(self.convc1(corr))))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        return torch.cat([F.relu(self.conv(torch.cat([cor, flo], 1))), flow], 1)


class SepConvGRU(nn.Module):
    def __init__(self, hidden, inp):
        super().__init__()
        for s, k, pad in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for g in "zrq":
                setattr(self, f"conv{g}{s}", nn.Conv2d(hidden + inp, hidden, k, padding=pad))

    def forward(self, h, x):
        for s in "12":
            hx = torch.cat([h, x], 1)
            z = torch.sigmoid(getattr(self, "convz" + s)(hx))
            r = torch.sigmoid(getattr(self, "convr" + s)(hx))
            q = torch.tanh(getattr(self, "convq" + s)(torch.cat([r * h, x], 1)))
            h = (1 - z) * h + z * q
        return h


class FlowHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(128, 256, 3, padding=1)
        self.conv2 = nn.Conv2d(256, 2, 3, padding=1)

    def forward(self, x):
        return self.conv2(F.relu(self.conv1(x)))


class UpdateBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = MotionEncoder(4 * 81)
        self.gru = SepConvGRU(128, 128 + 128)
        self.flow_head = FlowHead()
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(), nn.Conv2d(256, 64 * 9, 1))

    def forward(self, net, inp, corr, flow):
        net = self.gru(net, torch.cat([inp, self.encoder(flow, corr)], 1))
        return net, 0.25 * self.mask(net), self.flow_head(net)


class Upstream(nn.Module):
    def __init__(self):
        super().__init__()
        self.fnet = BasicEncoder(256, "instance")
        self.cnet = BasicEncoder(256, "batch")
        self.update_block = UpdateBlock()


@pytest.fixture(scope="module")
def converted(tmp_path_factory):
    torch.manual_seed(1234)
    m = Upstream().double().eval()
    with torch.no_grad():
        for mod in m.modules():  # non-trivial inference statistics and affine parameters
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_mean.uniform_(-0.5, 0.5)
                mod.running_var.uniform_(0.5, 2.0)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.3, 0.3)
    d = tmp_path_factory.mktemp("ckpt")
    pth, npz = str(d / "raft-things.pth"), str(d / "raft-things.npz")
    sd = {"module." + k: v.float() for k, v in m.state_dict().items()}  # checkpoints are fp32 DataParallel state dicts
    assert "module.cnet.layer2.0.norm3.running_var" in sd and "module.cnet.layer2.0.downsample.1.running_var" in sd
    assert "module.cnet.norm1.num_batches_tracked" in sd
    torch.save(sd, pth)
    assert convert.main([pth, npz]) == 0
    m = m.float().double()  # the fp32-rounded weights the checkpoint holds
    m.load_state_dict({k[len("module."):]: v.double() for k, v in sd.items()})
    p = {k: torch.from_numpy(v).double() for k, v in load_npz(npz).items()}
    return m.eval(), p


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def test_converted_keys_are_the_reference_names(converted):
    _, p = converted
    from raft_b200 import synth
    want = set(synth.make_weights(False))  # the reference's variable names for raft-things
    extra = {k for k in p if "/norm3/" in k}  # upstream registers the strided blocks' norm twice; the reference reads downsample.1
    assert want <= set(p), sorted(want - set(p))[:5]
    assert set(p) - want == extra, sorted(set(p) - want - extra)[:5]
    assert p["fnet/conv1/W"].shape == (7, 7, 3, 64) and p["update_block/gru/convz2/W"].shape == (5, 1, 384, 128)


@pytest.mark.parametrize("name,norm", [("fnet", "instance"), ("cnet", "batch")])
def test_encoder_matches_upstream_module(converted, name, norm):
    m, p = converted
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, 65, 97, generator=g, dtype=torch.float64) * 2 - 1  # 65 -> 33 -> 17 -> 9: odd at every strided conv
    with torch.no_grad():
        ref = _nhwc(getattr(m, name)(x))
    out = O.basic_encoder(_nhwc(x), p, name, norm)
    assert out.shape == ref.shape == (2, 9, 13, 256)
    assert (out - ref).abs().max().item() < 1e-9 * max(ref.abs().max().item(), 1.0)


def test_update_block_matches_upstream_module(converted):
    m, p = converted
    g = torch.Generator().manual_seed(8)
    B, h, w = 1, 9, 14
    net = torch.tanh(torch.randn(B, 128, h, w, generator=g, dtype=torch.float64))
    inp = torch.relu(torch.randn(B, 128, h, w, generator=g, dtype=torch.float64))
    corr = torch.randn(B, 324, h, w, generator=g, dtype=torch.float64)
    flow = torch.randn(B, 2, h, w, generator=g, dtype=torch.float64) * 3
    with torch.no_grad():
        rn, rm, rd = m.update_block(net, inp, corr, flow)
    on, om, od = O.basic_update_block(_nhwc(net), _nhwc(inp), _nhwc(corr), _nhwc(flow), p)
    for a, b in ((on, rn), (om, rm), (od, rd)):
        assert (a - _nhwc(b)).abs().max().item() < 1e-9 * max(b.abs().max().item(), 1.0)


# ---- raft-small: BottleneckBlock encoders (fnet: instance norm, cnet: no norm), ConvGRU update block --------------------
class BottleneckBlock(nn.Module):
    def __init__(self, cin, planes, norm, stride):
        super().__init__()
        q = planes // 4
        self.conv1 = nn.Conv2d(cin, q, 1)
        self.conv2 = nn.Conv2d(q, q, 3, padding=1, stride=stride)
        self.conv3 = nn.Conv2d(q, planes, 1)
        self.norm1, self.norm2, self.norm3 = _norm(norm, q), _norm(norm, q), _norm(norm, planes)
        self.downsample = None
        if stride != 1:
            self.norm4 = _norm(norm, planes)
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes, 1, stride=stride), self.norm4)

    def forward(self, x):
        y = F.relu(self.norm1(self.conv1(x)))
        y = F.relu(self.norm2(self.conv2(y)))
        y = F.relu(self.norm3(self.conv3(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return F.relu(x + y)


class SmallEncoder(nn.Module):
    def __init__(self, out_dim, norm):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 32, 7, stride=2, padding=3)
        self.norm1 = _norm(norm, 32)
        self.layer1 = nn.Sequential(BottleneckBlock(32, 32, norm, 1), BottleneckBlock(32, 32, norm, 1))
        self.layer2 = nn.Sequential(BottleneckBlock(32, 64, norm, 2), BottleneckBlock(64, 64, norm, 1))
        self.layer3 = nn.Sequential(BottleneckBlock(64, 96, norm, 2), BottleneckBlock(96, 96, norm, 1))
        self.conv2 = nn.Conv2d(96, out_dim, 1)

    def forward(self, x):
        x = F.relu(self.norm1(self.conv1(x)))
        return self.conv2(self.layer3(self.layer2(self.layer1(x))))


class SmallUpdateBlock(nn.Module):
    def __init__(self):
        super().__init__()
        enc = nn.Module()
        enc.convc1 = nn.Conv2d(4 * 49, 96, 1)
        enc.convf1 = nn.Conv2d(2, 64, 7, padding=3)
        enc.convf2 = nn.Conv2d(64, 32, 3, padding=1)
        enc.conv = nn.Conv2d(128, 80, 3, padding=1)
        self.encoder = enc
        gru = nn.Module()
        for g in "zrq":
            setattr(gru, "conv" + g, nn.Conv2d(96 + 82 + 64, 96, 3, padding=1))
        self.gru = gru
        fh = nn.Module()
        fh.conv1 = nn.Conv2d(96, 128, 3, padding=1)
        fh.conv2 = nn.Conv2d(128, 2, 3, padding=1)
        self.flow_head = fh

    def forward(self, net, inp, corr, flow):
        e, g, f = self.encoder, self.gru, self.flow_head
        cor = F.relu(e.convc1(corr))
        flo = F.relu(e.convf2(F.relu(e.convf1(flow))))
        x = torch.cat([inp, F.relu(e.conv(torch.cat([cor, flo], 1))), flow], 1)
        hx = torch.cat([net, x], 1)
        z, r = torch.sigmoid(g.convz(hx)), torch.sigmoid(g.convr(hx))
        q = torch.tanh(g.convq(torch.cat([r * net, x], 1)))
        net = (1 - z) * net + z * q
        return net, f.conv2(F.relu(f.conv1(net)))


class UpstreamSmall(nn.Module):
    def __init__(self):
        super().__init__()
        self.fnet = SmallEncoder(128, "instance")
        self.cnet = SmallEncoder(96 + 64, "none")
        self.update_block = SmallUpdateBlock()


@pytest.fixture(scope="module")
def converted_small(tmp_path_factory):
    torch.manual_seed(4321)
    m = UpstreamSmall().eval()
    d = tmp_path_factory.mktemp("ckpt_small")
    pth, npz = str(d / "raft-small.pth"), str(d / "raft-small.npz")
    torch.save({"module." + k: v for k, v in m.state_dict().items()}, pth)
    assert convert.main([pth, npz]) == 0
    return m.double().eval(), {k: torch.from_numpy(v).double() for k, v in load_npz(npz).items()}


def test_small_model_matches_upstream_modules(converted_small):
    m, p = converted_small
    from raft_b200 import synth
    assert set(p) == set(synth.make_weights(True))  # exactly the reference's variable names for raft-small
    g = torch.Generator().manual_seed(11)
    x = torch.rand(1, 3, 73, 57, generator=g, dtype=torch.float64) * 2 - 1  # 73 -> 37 -> 19 -> 10, 57 -> 29 -> 15 -> 8
    with torch.no_grad():
        for name, norm in (("fnet", "instance"), ("cnet", "none")):
            ref = _nhwc(getattr(m, name)(x))
            out = O.small_encoder(_nhwc(x), p, name, norm)
            assert out.shape == ref.shape and (out - ref).abs().max().item() < 1e-9 * max(ref.abs().max().item(), 1.0), name
        B, h, w = 1, 8, 11
        net = torch.tanh(torch.randn(B, 96, h, w, generator=g, dtype=torch.float64))
        inp = torch.relu(torch.randn(B, 64, h, w, generator=g, dtype=torch.float64))
        corr = torch.randn(B, 196, h, w, generator=g, dtype=torch.float64)
        flow = torch.randn(B, 2, h, w, generator=g, dtype=torch.float64) * 3
        rn, rd = m.update_block(net, inp, corr, flow)
    on, om, od = O.small_update_block(_nhwc(net), _nhwc(inp), _nhwc(corr), _nhwc(flow), p)
    assert om is None
    for a, b in ((on, rn), (od, rd)):
        assert (a - _nhwc(b)).abs().max().item() < 1e-9 * max(b.abs().max().item(), 1.0)



# ---- correlation pyramid + lookup against an upstream-style CorrBlock on torch.grid_sample -------------------------------
def _upstream_corr_lookup(fmap1, fmap2, coords, radius, levels=4):
    """All-pairs volume / sqrt(C), 2x2 average pyramid over the target axes, (2r+1)^2 bilinear taps per level around
    coords / 2^i with the FIRST window axis added to x (the published implementation stacks meshgrid(dy, dx) onto (x, y)
    coordinates); zero outside.  fmap NCHW, coords [B,2,H,W] (x, y) -> [B, levels*(2r+1)^2, H, W]."""
    B, C, H, W = fmap1.shape
    corr = torch.matmul(fmap1.view(B, C, H * W).transpose(1, 2), fmap2.view(B, C, H * W)) / C ** 0.5
    corr = corr.view(B * H * W, 1, H, W)
    pyr = [corr]
    for _ in range(levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    c = coords.permute(0, 2, 3, 1).reshape(B * H * W, 1, 1, 2)
    d = torch.linspace(-radius, radius, 2 * radius + 1, dtype=coords.dtype)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1)  # delta[i, j] = (d[i], d[j]) added to (x, y)
    out = []
    for i, vol in enumerate(pyr):
        xy = c / 2 ** i + delta[None]
        hl, wl = vol.shape[-2:]
        grid = torch.stack([2 * xy[..., 0] / (wl - 1) - 1, 2 * xy[..., 1] / (hl - 1) - 1], dim=-1)
        s = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        out.append(s.view(B, H, W, -1))
    return torch.cat(out, dim=-1)


@pytest.mark.parametrize("radius", [4, 3])
def test_pyramid_lookup_matches_grid_sample_corr_block(radius):
    """Volume scaling, VALID 2x2 pooling, coords / 2^i, x-major window and level order of the oracle (= the reference's
    GetCorrPyramid / SampleCorr) against torch.grid_sample, on the taps that lie strictly inside their level (at and beyond
    the border the reference's truncating / clamping sampler deliberately differs from zero padding: utils.py:54-89)."""
    g = torch.Generator().manual_seed(21)
    B, C, H, W = 1, 16, 40, 72
    f1 = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    f2 = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    coords = torch.stack([xs, ys], 0)[None] + (torch.rand(B, 2, H, W, generator=g, dtype=torch.float64) * 6 - 3)
    ref = _upstream_corr_lookup(f1, f2, coords, radius)
    pyr = O.get_corr_pyramid(_nhwc(f1), _nhwc(f2))
    out = O.sample_corr(pyr, _nhwc(coords), radius=radius)
    assert out.shape == ref.shape == (B, H, W, 4 * (2 * radius + 1) ** 2)
    n = (2 * radius + 1) ** 2
    d = torch.linspace(-radius, radius, 2 * radius + 1, dtype=torch.float64)
    dx, dy = d[:, None].expand(-1, 2 * radius + 1).reshape(-1), d[None, :].expand(2 * radius + 1, -1).reshape(-1)
    checked = 0
    for i in range(4):
        hl, wl = H >> i, W >> i
        x = _nhwc(coords)[..., 0:1] / 2 ** i + dx
        y = _nhwc(coords)[..., 1:2] / 2 ** i + dy
        inside = (x > 0) & (x < wl - 1) & (y > 0) & (y < hl - 1)
        a, b = out[..., i * n:(i + 1) * n][inside], ref[..., i * n:(i + 1) * n][inside]
        assert inside.float().mean().item() > (0.5 if i < 3 else 0.1), i
        assert (a - b).abs().max().item() < 1e-10 * max(b.abs().max().item(), 1.0), f"level {i}"
        checked += int(inside.sum())
    assert checked > 100000


# ---- the output edge: convex upsampling and upflow8 against their torch forms ---------------------------------------------
def test_convex_upsampling_matches_unfold_form():
    """RAFT.upsample_flow (RAFT.py:118-134: softmax over the 9 taps, tf.extract_image_patches of 8*flow, reshape (9, 2))
    against the published torch form (softmax over dim 2 of mask.view(N,1,9,8,8,H,W), F.unfold of 8*flow): pins the
    (ky, kx, c) depth order of extract_image_patches and the (9, 8, 8) layout of the 576 mask channels independently."""
    g = torch.Generator().manual_seed(31)
    N, H, W = 2, 5, 7
    flow = torch.randn(N, 2, H, W, generator=g, dtype=torch.float64) * 3
    mask = torch.randn(N, 576, H, W, generator=g, dtype=torch.float64)
    m = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    ref = torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(N, 2, 8 * H, 8 * W)
    out = O.upsample_flow(_nhwc(flow), _nhwc(mask))
    assert out.shape == (N, 8 * H, 8 * W, 2)
    assert (out - _nhwc(ref)).abs().max().item() < 1e-12 * max(ref.abs().max().item(), 1.0)


def test_upflow8_is_align_corners_resize_without_the_factor_8():
    """utils.py:105-111: tf.image.resize_bilinear(align_corners=True) -- and, unlike upstream, NOT multiplied by 8."""
    g = torch.Generator().manual_seed(32)
    flow = torch.randn(2, 2, 6, 9, generator=g, dtype=torch.float64)
    ref = F.interpolate(flow, size=(48, 72), mode="bilinear", align_corners=True)
    out = O.upflow8(_nhwc(flow))
    # the oracle forms the source coordinates in float32 like TF's kernel (out_idx * float((in-1)/(out-1))): ~1e-6 of a pixel
    assert (out - _nhwc(ref)).abs().max().item() < 1e-5
