"""F4 end to end on the CPU: a checkpoint in the UPSTREAM (PyTorch RAFT) layout -- real ``torch.nn`` modules with the
published architecture, their genuine ``state_dict()`` keys ('module.' prefix of DataParallel, BatchNorm bookkeeping entries,
the norm of a strided block registered twice as ``norm3`` and ``downsample.1``) saved with ``torch.save`` -- goes through the
converter CLI into the reference's ``.npz`` naming, and the oracle (the restatement of the reference's TF graph) must then
compute what the torch modules compute.  That pins the key mapping, OIHW -> HWIO, the BatchNorm statistics and, as a side
effect, the oracle's block structure against an independent implementation.

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
