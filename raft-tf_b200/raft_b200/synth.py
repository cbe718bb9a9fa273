"""Synthetic inputs and seeded random weights (reference npz key scheme).

There is no network access, so neither Sintel/KITTI frames nor the reference's pretrained
``release_weight/*.npz`` (readme.md:28, git-ignored) exist here.  ``make_pair`` builds smooth
textured frame pairs related by a smooth flow; ``make_weights`` builds an
``.npz``-style dict keyed exactly like the TF variables the reference creates
(``<scope>/W`` HWIO, ``<scope>/b``, BN ``gamma``/``beta``/``mean/EMA``/``variance/EMA``).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def _blur(a: np.ndarray, sigma: float) -> np.ndarray:
    from scipy.ndimage import gaussian_filter
    return gaussian_filter(a, sigma=(sigma, sigma, 0), mode="wrap")


def make_pair(H: int, W: int, seed: int = 1000, amp: float = 12.0) -> Tuple[np.ndarray, np.ndarray]:
    """One synthetic frame pair, float32 [H,W,3] in [0,1]."""
    from scipy.ndimage import map_coordinates
    rng = np.random.default_rng(seed)
    base = rng.random((H + 64, W + 64, 3), dtype=np.float32)
    base = 0.6 * _blur(base, 3.0) + 0.4 * _blur(base, 1.0)
    base -= base.min()
    base /= max(base.max(), 1e-6)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    ph = rng.uniform(0, 2 * np.pi, size=6)
    u = amp * (0.5 * np.sin(2 * np.pi * xs / W + ph[0]) + 0.3 * np.sin(2 * np.pi * ys / H * 2 + ph[1])
               + 0.2 * np.sin(2 * np.pi * (xs + ys) / (W + H) * 3 + ph[2]))
    v = amp * 0.5 * (0.5 * np.sin(2 * np.pi * ys / H + ph[3]) + 0.3 * np.sin(2 * np.pi * xs / W * 2 + ph[4])
                     + 0.2 * np.sin(2 * np.pi * (xs - ys) / (W + H) * 3 + ph[5]))
    f1 = base[32:32 + H, 32:32 + W].copy()
    f2 = np.stack([map_coordinates(base[..., c], [ys + 32 - v, xs + 32 - u], order=1, mode="nearest")
                   for c in range(3)], axis=-1)
    f2 = f2 + rng.normal(0, 0.01, f2.shape).astype(np.float32)
    return f1.astype(np.float32), np.clip(f2, 0, 1).astype(np.float32)


def make_batch(B: int, H: int, W: int, seed0: int = 1000):
    pairs = [make_pair(H, W, seed0 + i) for i in range(B)]
    return np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])


# ------------------------------------------------------------------------------------
def _encoder_shapes(name: str, small: bool, out_dim: int, norm_fn: str):
    convs, norms = {}, []
    if not small:  # BasicEncoder, model_utils.py:61-82
        convs[f"{name}/conv1"] = (7, 7, 3, 64)
        norms.append((f"{name}/norm1", 64))
        cin = 64
        for lname, dim, stride in (("layer1", 64, 1), ("layer2", 96, 2), ("layer3", 128, 2)):
            for blk, (ci, st) in enumerate(((cin, stride), (dim, 1))):
                s = f"{name}/{lname}/{blk}"
                convs[s + "/conv1"] = (3, 3, ci, dim)
                convs[s + "/conv2"] = (3, 3, dim, dim)
                norms += [(s + "/norm1", dim), (s + "/norm2", dim)]
                if st != 1:
                    convs[s + "/downsample.0"] = (1, 1, ci, dim)
                    norms.append((s + "/downsample.1", dim))
            cin = dim
        convs[f"{name}/conv2"] = (1, 1, 128, out_dim)
    else:  # SmallEncoder, model_utils.py:84-105
        convs[f"{name}/conv1"] = (7, 7, 3, 32)
        norms.append((f"{name}/norm1", 32))
        cin = 32
        for lname, dim, stride in (("layer1", 32, 1), ("layer2", 64, 2), ("layer3", 96, 2)):
            for blk, (ci, st) in enumerate(((cin, stride), (dim, 1))):
                s = f"{name}/{lname}/{blk}"
                convs[s + "/conv1"] = (1, 1, ci, dim // 4)
                convs[s + "/conv2"] = (3, 3, dim // 4, dim // 4)
                convs[s + "/conv3"] = (1, 1, dim // 4, dim)
                norms += [(s + "/norm1", dim // 4), (s + "/norm2", dim // 4), (s + "/norm3", dim)]
                if st != 1:
                    convs[s + "/downsample.0"] = (1, 1, ci, dim)
                    norms.append((s + "/downsample.1", dim))
            cin = dim
        convs[f"{name}/conv2"] = (1, 1, 96, out_dim)
    return convs, (norms if norm_fn == "batch" else [])


def update_block_shapes(small: bool) -> Dict[str, tuple]:
    """Hot-path conv shapes (SURVEY 8(a) 'Hot-path weights')."""
    u = "update_block"
    if not small:
        return {
            f"{u}/encoder/convc1": (1, 1, 324, 256), f"{u}/encoder/convc2": (3, 3, 256, 192),
            f"{u}/encoder/convf1": (7, 7, 2, 128), f"{u}/encoder/convf2": (3, 3, 128, 64),
            f"{u}/encoder/conv": (3, 3, 256, 126),
            f"{u}/gru/convz1": (1, 5, 384, 128), f"{u}/gru/convr1": (1, 5, 384, 128),
            f"{u}/gru/convq1": (1, 5, 384, 128),
            f"{u}/gru/convz2": (5, 1, 384, 128), f"{u}/gru/convr2": (5, 1, 384, 128),
            f"{u}/gru/convq2": (5, 1, 384, 128),
            f"{u}/flow_head/conv1": (3, 3, 128, 256), f"{u}/flow_head/conv2": (3, 3, 256, 2),
            f"{u}/mask/0": (3, 3, 128, 256), f"{u}/mask/2": (1, 1, 256, 576),
        }
    return {
        f"{u}/encoder/convc1": (1, 1, 196, 96), f"{u}/encoder/convf1": (7, 7, 2, 64),
        f"{u}/encoder/convf2": (3, 3, 64, 32), f"{u}/encoder/conv": (3, 3, 128, 80),
        f"{u}/gru/convz": (3, 3, 242, 96), f"{u}/gru/convr": (3, 3, 242, 96),
        f"{u}/gru/convq": (3, 3, 242, 96),
        f"{u}/flow_head/conv1": (3, 3, 96, 128), f"{u}/flow_head/conv2": (3, 3, 128, 2),
    }


def make_weights(small: bool = False, seed: int = 7, gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded random weights in the reference's variable naming.

    He-normal conv kernels (std = gain*sqrt(2/fan_in)), small non-zero biases (so a dropped
    bias is caught), ``flow_head/conv2`` scaled by 0.05 and ``mask/2`` by 0.5 so that the
    recurrence moves the flow by a fraction of a pixel per iteration instead of exploding;
    BN statistics are non-trivial (gamma~U(.8,1.2), beta~N(0,.05), mean~N(0,.05), var~U(.8,1.2)).
    """
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}
    convs: Dict[str, tuple] = {}
    norms = []
    hidden, ctx, fdim = (96, 64, 128) if small else (128, 128, 256)
    c, n = _encoder_shapes("fnet", small, fdim, "instance")
    convs.update(c)
    c, n2 = _encoder_shapes("cnet", small, hidden + ctx, "none" if small else "batch")
    convs.update(c)
    norms += n + n2
    convs.update(update_block_shapes(small))
    for k, shp in convs.items():
        fan_in = shp[0] * shp[1] * shp[2]
        std = gain * np.sqrt(2.0 / fan_in)
        if k.endswith("flow_head/conv2"):
            std *= 0.05
        if k.endswith("mask/2"):
            std *= 0.5
        p[k + "/W"] = rng.normal(0, std, shp).astype(np.float32)
        p[k + "/b"] = rng.normal(0, 0.02, (shp[3],)).astype(np.float32)
    for s, ch in norms:
        p[s + "/gamma"] = rng.uniform(0.8, 1.2, ch).astype(np.float32)
        p[s + "/beta"] = rng.normal(0, 0.05, ch).astype(np.float32)
        p[s + "/mean/EMA"] = rng.normal(0, 0.05, ch).astype(np.float32)
        p[s + "/variance/EMA"] = rng.uniform(0.8, 1.2, ch).astype(np.float32)
    return p
