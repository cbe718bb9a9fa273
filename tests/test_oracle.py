"""CPU tests of the oracle itself: every reference quirk in SURVEY 8(c) is a named test."""
import numpy as np
import pytest
import torch

from oracle import raft_oracle as O


def test_sampler_border_table():
    """W=6 image 5,15,..,55; x<=-1 -> img[0]; -1<x<0 extrapolates; x>=W-1 -> img[W-1]."""
    img = torch.tensor([5., 15, 25, 35, 45, 55]).reshape(1, 1, 6, 1)
    xs = torch.tensor([-3.0, -1.0, -0.75, -0.25, 0.0, 0.5, 2.25, 4.999, 5.0, 7.3])
    coords = torch.stack([xs, torch.zeros_like(xs)], -1).reshape(1, 1, -1, 2)
    out = O.tf_grid_sample(img, coords).flatten()
    exp = [5.0, 5.0, (1 + 0.75) * 5 - 0.75 * 15, (1 + 0.25) * 5 - 0.25 * 15, 5.0, 10.0, 27.5, 54.99, 55.0, 55.0]
    assert torch.allclose(out, torch.tensor(exp), atol=1e-4)
    assert abs(out[2].item() - (-2.5)) < 1e-6  # the value quoted in SURVEY


def test_trunc_not_floor():
    img = torch.arange(12.).reshape(1, 3, 4, 1)
    c = torch.tensor([[-0.5, -0.5]]).reshape(1, 1, 1, 2)
    # trunc -> x0=y0=0, x1=y1=1; qx=qy=1.5 -> weights 2.25,-0.75,-0.75,0.25
    v = O.tf_grid_sample(img, c).item()
    assert abs(v - (2.25 * 0 - 0.75 * 4 - 0.75 * 1 + 0.25 * 5)) < 1e-6


def test_coords_grid_xy():
    g = O.coords_grid(2, 3, 5)
    assert g.shape == (2, 3, 5, 2)
    assert g[1, 2, 4, 0] == 4 and g[1, 2, 4, 1] == 2


def test_same_padding_stride2():
    assert O._same_pad(432, 7, 2) == (2, 3)
    assert O._same_pad(216, 3, 2) == (0, 1)
    assert O._same_pad(216, 1, 2) == (0, 0)
    assert O._same_pad(55, 3, 1) == (1, 1)
    assert O._same_pad(55, 5, 1) == (2, 2)


def test_window_order_x_major_and_centre_tap():
    """Lookup at integer coords: tap k=(dx+r)(2r+1)+(dy+r) reads corr[n, y+dy, x+dx]."""
    torch.manual_seed(0)
    b, h, w, r = 1, 8, 16, 3
    vol = torch.randn(b * h * w, h, w, 1)
    pyr = [vol] + [torch.zeros(b * h * w, h >> l, w >> l, 1) for l in (1, 2, 3)]
    coords = O.coords_grid(b, h, w)
    out = O.sample_corr(pyr, coords, radius=r)
    K = (2 * r + 1) ** 2
    n = 3 * w + 5  # pixel (y=3, x=5)
    centre = r * (2 * r + 1) + r
    assert out[0, 3, 5, centre] == vol[n, 3, 5, 0]
    k = (2 + r) * (2 * r + 1) + (-1 + r)  # dx=+2, dy=-1
    assert out[0, 3, 5, k] == vol[n, 2, 7, 0]
    assert out.shape[-1] == 4 * K


def test_pool_linearity_valid_floor():
    """pooled volume == volume of pooled fmap2 (odd dims, VALID floor) -- used by the TC corr build."""
    torch.manual_seed(1)
    f1 = torch.randn(1, 13, 27, 16, dtype=torch.float64)
    f2 = torch.randn(1, 13, 27, 16, dtype=torch.float64)
    pyr = O.get_corr_pyramid(f1, f2)
    assert [tuple(p.shape[1:3]) for p in pyr] == [(13, 27), (6, 13), (3, 6), (1, 3)]
    f2p = f2
    for l in (1, 2, 3):
        f2p = O.avg_pool_2x2_valid(f2p)
        v = torch.matmul(f1.reshape(1, 13 * 27, 16), f2p.reshape(1, -1, 16).transpose(1, 2)) / 4.0
        assert torch.allclose(pyr[l].reshape(1, 13 * 27, -1), v, atol=1e-12)


def test_upflow8_no_times8_and_align_corners():
    flow = torch.zeros(1, 2, 3, 2)
    flow[0, :, :, 0] = torch.tensor([[0., 1, 2], [3, 4, 5]])
    up = O.upflow8(flow)
    assert up.shape == (1, 16, 24, 2)
    assert up[0, 0, 0, 0] == 0 and abs(up[0, -1, -1, 0].item() - 5.0) < 1e-6  # corners map to corners, no x8
    assert abs(up[0, 0, -1, 0].item() - 2.0) < 1e-6


def test_convex_upsample_matches_unfold_formulation():
    """RAFT.py:119-134 against the upstream unfold formulation (fp64)."""
    torch.manual_seed(2)
    b, h, w = 1, 3, 4
    flow = torch.randn(b, h, w, 2, dtype=torch.float64)
    mask = torch.randn(b, h, w, 576, dtype=torch.float64)
    up = O.upsample_flow(flow, mask)
    m = mask.permute(0, 3, 1, 2).reshape(b, 1, 9, 8, 8, h, w).softmax(2)
    uf = torch.nn.functional.unfold(8 * flow.permute(0, 3, 1, 2), [3, 3], padding=1).reshape(b, 2, 9, 1, 1, h, w)
    ref = (m * uf).sum(2).permute(0, 1, 4, 2, 5, 3).reshape(b, 2, 8 * h, 8 * w).permute(0, 2, 3, 1)
    assert torch.allclose(up, ref, atol=1e-12)


def test_split_format_precision():
    """hi + lo*2^-11 reproduces fp32 to ~2^-22 relative (the TC operand format)."""
    x = torch.randn(100000) * 3
    hi = x.half().float()
    lo = ((x - hi) * 2048).half().float()
    err = ((hi + lo / 2048) - x).abs() / x.abs().clamp_min(1e-3)
    assert err.max() < 2 ** -21


@pytest.mark.parametrize("small", [False, True])
def test_oracle_forward_shapes(small):
    from raft_b200 import synth
    w = synth.make_weights(small)
    l, r = synth.make_batch(1, 64, 96)
    m = O.RAFTOracle(w, small=small, iters=2)
    up, low = m.forward(torch.from_numpy(l), torch.from_numpy(r), return_lowres=True)
    assert up.shape == (1, 64, 96, 2) and low.shape == (1, 8, 12, 2)
    assert torch.isfinite(up).all()
