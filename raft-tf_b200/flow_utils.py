"""Flow colour coding used by the CLI (the reference's infer_raft.py:25,43 calls
flow_utils.flow_to_color(flow, convert_to_bgr=True)).  Independent numpy implementation of the
Middlebury colour wheel (Baker et al., ICCV 2007): same normalisation by the maximum radius
(+1e-5) and floor(255*col) quantisation as the reference's flow_utils.py:95-121."""
import numpy as np

_SEGMENTS = (("RY", 15), ("YG", 6), ("GC", 4), ("CB", 11), ("BM", 13), ("MR", 6))


def make_colorwheel():
    n = sum(k for _, k in _SEGMENTS)
    wheel = np.zeros((n, 3))
    # each segment ramps one channel up or down while another stays saturated
    plan = [(0, 1, +1), (1, 0, -1), (1, 2, +1), (2, 1, -1), (2, 0, +1), (0, 2, -1)]
    pos = 0
    for (_, k), (sat, ramp, sign) in zip(_SEGMENTS, plan):
        t = np.floor(255 * np.arange(k) / k)
        wheel[pos:pos + k, sat] = 255
        wheel[pos:pos + k, ramp] = t if sign > 0 else 255 - t
        pos += k
    return wheel


def flow_to_color(flow_uv, clip_flow=None, convert_to_bgr=False):
    """[H,W,2] flow -> [H,W,3] uint8 colour image.  Pinned bit for bit against the reference function run in-process
    (oracle/make_flowcolor_golden.py -> tests/golden/flow_to_color.npz), which fixes the floating-point TYPE of every
    step for float32 flows (flow_utils.py:58-121): direction / radius in the flow's dtype, interpolation weight and
    colours in float64, floor(255*col) quantisation."""
    flow_uv = np.asarray(flow_uv)
    assert flow_uv.ndim == 3 and flow_uv.shape[2] == 2
    if clip_flow is not None:
        flow_uv = np.clip(flow_uv, 0, clip_flow)
    u, v = flow_uv[..., 0], flow_uv[..., 1]
    ft = u.dtype if u.dtype.kind == "f" else np.dtype(np.float64)
    rad = np.sqrt(np.square(u) + np.square(v))
    den = (rad.max() + np.asarray(1e-5, dtype=ft)).astype(ft)  # rad_max + epsilon in the flow's precision
    u, v = (u / den).astype(ft), (v / den).astype(ft)
    wheel = make_colorwheel()
    n = wheel.shape[0]
    rad = np.sqrt(np.square(u) + np.square(v))
    fk = ((np.arctan2(-v, -u) / np.asarray(np.pi, dtype=ft)).astype(ft) + 1) / 2 * (n - 1) + 1
    k0 = np.minimum(np.floor(fk).astype(np.int32), n - 2)
    k1 = k0 + 1
    k1[k1 == n] = 1
    f = fk.astype(np.float64) - k0  # float64 from here on, like (float32 array - int32 array) in numpy
    inside = rad <= 1
    img = np.zeros(u.shape + (3,), np.uint8)
    for ch in range(3):
        col = (1 - f) * (wheel[k0, ch] / 255.0) + f * (wheel[k1, ch] / 255.0)
        col = np.where(inside, 1 - rad * (1 - col), col * 0.75)
        img[..., 2 - ch if convert_to_bgr else ch] = np.floor(255 * col)
    return img


def write_flo(path, flow_uv):
    """Middlebury .flo file (magic 202021.25, int32 width, int32 height, float32 [H,W,2] row-major u,v) -- the
    format of the reference's unused writer (flow_utils.py:302-318)."""
    f = np.ascontiguousarray(flow_uv, dtype=np.float32)
    assert f.ndim == 3 and f.shape[2] == 2
    with open(path, "wb") as fh:
        np.array([202021.25], np.float32).tofile(fh)
        np.array([f.shape[1], f.shape[0]], np.int32).tofile(fh)
        f.tofile(fh)


def read_flo(path):
    with open(path, "rb") as fh:
        magic = np.fromfile(fh, np.float32, 1)[0]
        assert abs(magic - 202021.25) < 1e-3, "not a .flo file"
        w, h = np.fromfile(fh, np.int32, 2)
        return np.fromfile(fh, np.float32, 2 * int(w) * int(h)).reshape(int(h), int(w), 2)
