"""Parity properties at BASELINE.json's FULL sizes, where the CPU oracle would take too long for a plain
comparison (configs[1]: 1x440x1024 things/32 it; configs[2]: 8x544x960; configs[4]: small 768x1024).
Size-independent properties of the path are checked instead:
  * lookup at integer coordinates == direct indexing of the pyramid (bit exact), centre tap == vol[n, y, x];
  * pooled level l == 2x2 VALID average pool of level l-1 (fp32 rounding of the linearity trick only);
  * a batched forward equals the per-sample forwards bit for bit (samples never mix -> the batch shard of
    SURVEY 8(e) is exact), and is deterministic;
  * the full-size single-pair forward agrees with the oracle on a CROP-INDEPENDENT statistic: the oracle is run on
    the full frame once (a few seconds) for config[1] only."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import raft_oracle as O

pytestmark = pytest.mark.gpu


def test_lookup_integer_coords_equals_direct_index_fullsize(cuda):
    from networks.model_utils import GetCorrPyramid, SampleCorr
    B, h, w, C, r = 1, 55, 128, 256, 4
    g = torch.Generator(device="cpu").manual_seed(0)
    f1 = torch.randn(B, h, w, C, generator=g).to(cuda)
    f2 = torch.randn(B, h, w, C, generator=g).to(cuda)
    pyr = GetCorrPyramid(f1, f2)
    grid = O.coords_grid(B, h, w).to(cuda)
    out = SampleCorr(pyr, grid, radius=r)  # coords = grid -> every tap sits on an integer position
    K, D = (2 * r + 1) ** 2, 2 * r + 1
    vol = pyr[0].view(B, h, w, h, w)
    centre = out[..., r * D + r]
    ys, xs = torch.meshgrid(torch.arange(h, device=cuda), torch.arange(w, device=cuda), indexing="ij")
    assert torch.equal(centre[0], vol[0, ys, xs, ys, xs])
    # tap (dx=+2, dy=-1) of level 0 reads vol[n, y-1, x+2] wherever that position is inside the image (outside, the
    # reference's clamped-x1 weights give e.g. -2*I + 3*I, equal to I only up to fp32 rounding)
    k = (2 + r) * D + (-1 + r)
    inside = (ys - 1 >= 0) & (xs + 2 <= w - 1)
    yy, xx = (ys - 1).clamp(0, h - 1), (xs + 2).clamp(0, w - 1)
    assert torch.equal(out[0, ..., k][inside], vol[0, ys, xs, yy, xx][inside])
    assert torch.allclose(out[0, ..., k], vol[0, ys, xs, yy, xx], rtol=1e-6, atol=1e-6)
    # level 1 centre tap at coords/2: integer for even coordinates
    l1 = pyr[1].view(B, h, w, h // 2, w // 2)
    ev = out[0, 0::2, 0::2, K + r * D + r]
    assert torch.equal(ev[: (h // 2), :], l1[0, ys[0::2, 0::2][: h // 2], xs[0::2, 0::2][: h // 2],
                                             (ys[0::2, 0::2] // 2)[: h // 2], (xs[0::2, 0::2] // 2)[: h // 2]])


def test_pyramid_levels_are_valid_avgpools_fullsize(cuda):
    from networks.model_utils import GetCorrPyramid
    B, h, w, C = 2, 68, 120, 256  # KITTI grid of configs[2] (odd level dims: 34x60, 17x30, 8x15)
    g = torch.Generator(device="cpu").manual_seed(1)
    f1 = torch.randn(B, h, w, C, generator=g).to(cuda)
    f2 = torch.randn(B, h, w, C, generator=g).to(cuda)
    pyr = GetCorrPyramid(f1, f2)
    scale = pyr[0].abs().max().item()
    for l in range(1, 4):
        ref = torch.nn.functional.avg_pool2d(pyr[l - 1].permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        assert pyr[l].shape == ref.shape
        assert (pyr[l] - ref).abs().max().item() < 2e-5 * scale, l


@pytest.mark.parametrize("small,B,H,W,iters", [(False, 3, 272, 480, 4), (True, 4, 384, 512, 3)])
def test_batched_equals_per_sample_bit_exact(cuda, small, B, H, W, iters):
    from raft_b200 import synth
    from networks.RAFT import RAFT
    p = synth.make_weights(small)
    l, r = synth.make_batch(B, H, W)
    mb = RAFT((H, W, 3), SimpleNamespace(small=small), iters=iters, batch=B, device=cuda).load(p)
    full = mb.forward(l, r).clone()
    again = mb.forward(l, r).clone()
    assert torch.equal(full, again), "forward is not deterministic"
    m1 = RAFT((H, W, 3), SimpleNamespace(small=small), iters=iters, batch=1, device=cuda).load(p)
    for i in range(B):
        one = m1.forward(l[i:i + 1], r[i:i + 1])
        assert torch.equal(one[0], full[i]), f"sample {i} differs between batched and single execution"


def test_config1_fullsize_matches_oracle(cuda):
    """BASELINE configs[1] at full size (436x1024 -> 440x1024) but 4 iterations to keep the CPU oracle at a few
    seconds; the 32-iteration error growth is covered at 128x256 in test_gpu_e2e.py."""
    from raft_b200 import synth
    from networks.RAFT import RAFT
    p = synth.make_weights(False)
    l, r = synth.make_batch(1, 436, 1024)
    lp = np.pad(l, ((0, 0), (2, 2), (0, 0), (0, 0)), mode="edge")
    rp = np.pad(r, ((0, 0), (2, 2), (0, 0), (0, 0)), mode="edge")
    ref = O.RAFTOracle(p, iters=4).forward(torch.from_numpy(lp), torch.from_numpy(rp))[:, 2:-2]
    out = RAFT((436, 1024, 3), SimpleNamespace(small=False), iters=4, device=cuda).load(p).forward(l, r).cpu()
    assert out.shape == (1, 436, 1024, 2)
    e = (out - ref).abs().max().item()
    assert e < 1e-3, e
