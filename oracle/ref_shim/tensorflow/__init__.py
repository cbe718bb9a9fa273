"""Minimal eager numpy stand-in for the subset of the TF1 API that gonglixue/RAFT-tf's networks/*.py call.
TEST INFRASTRUCTURE (see ../README.md).  float32 throughout; semantics follow the TF documentation."""
import contextlib

import numpy as np

__version__ = "1.15.0-numpy-shim"
float32, int32 = np.float32, np.int32
AUTO_REUSE = "AUTO_REUSE"


class TShape(tuple):
    def as_list(self):
        return [int(v) for v in self]


class T(np.ndarray):
    """ndarray that answers the few TF tensor methods the reference uses."""

    def __new__(cls, a):
        return np.asarray(a).view(cls)

    @property
    def shape(self):
        return TShape(np.ndarray.shape.__get__(self))

    def get_shape(self):
        return self.shape


def _t(x):
    return T(x) if isinstance(x, np.ndarray) else x


def _shape_arg(s):
    if isinstance(s, np.ndarray):
        s = s.tolist()
    return tuple(int(v) for v in s)


# ---- scopes (only variable_scope contributes to variable names) ---------------------------------
_scope = []


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


@contextlib.contextmanager
def name_scope(name):
    yield


def current_scope(name=None):
    parts = list(_scope) + ([name] if name else [])
    return "/".join(parts)


VARIABLES = {}  # reference variable name -> np.ndarray (float32); set by the caller


def get_variable(name, shape=None, initializer=None, trainable=None):
    return T(VARIABLES[current_scope(name)])


def constant_initializer(value=0.0):
    return ("const", value)


# ---- tensor construction / shape ------------------------------------------------------------------
def shape(x):
    return T(np.array(np.ndarray.shape.__get__(np.asarray(x)), dtype=np.int32))


def unstack(x):
    return [x[i] for i in range(len(x))]


def stack(values, axis=0):
    return T(np.stack([np.asarray(v) for v in values], axis=axis))


def reshape(x, shp, name=None):
    return T(np.asarray(x).reshape(_shape_arg(shp)))


def cast(x, dtype):
    a = np.asarray(x)
    if np.issubdtype(dtype, np.integer) and np.issubdtype(a.dtype, np.floating):
        a = np.trunc(a)  # C++ static_cast<int>: toward zero
    return T(a.astype(dtype))


def range(*args):  # noqa: A001
    return T(np.arange(*[int(a) for a in args], dtype=np.int32))


def meshgrid(*xs):
    return [T(a) for a in np.meshgrid(*[np.asarray(x) for x in xs])]  # default indexing='xy' like TF


def linspace(start, stop, num):
    return T(np.linspace(start, stop, int(num), dtype=np.float32))


def expand_dims(x, axis):
    return T(np.expand_dims(np.asarray(x), axis))


def repeat(x, repeats, axis):
    return T(np.repeat(np.asarray(x), int(repeats), axis=axis))


def tile(x, multiples):
    return T(np.tile(np.asarray(x), _shape_arg(multiples)))


def zeros(shp, dtype=np.float32):
    return T(np.zeros(_shape_arg(shp), dtype=dtype))


def constant(v, dtype=None):
    return T(np.array(v, dtype=dtype))


def identity(x, name=None):
    return x


def stop_gradient(x):
    return x


def concat(values, axis, name=None):
    return T(np.concatenate([np.asarray(v) for v in values], axis=axis))


def split(x, sizes, axis):
    idx = np.cumsum(sizes)[:-1]
    return [T(a) for a in np.split(np.asarray(x), idx, axis=axis)]


def transpose(x, perm):
    return T(np.transpose(np.asarray(x), perm))


# ---- math -------------------------------------------------------------------------------------------
def clip_by_value(x, lo, hi):
    return T(np.clip(np.asarray(x), lo, hi))


def gather_nd(params, indices):
    idx = np.asarray(indices)
    return T(np.asarray(params)[tuple(idx[..., i] for i in np.arange(idx.shape[-1]))])


def add_n(xs):
    out = xs[0]
    for x in xs[1:]:
        out = out + x
    return out


def matmul(a, b):
    return T(np.matmul(np.asarray(a), np.asarray(b)))


def divide(a, b):
    return T(np.asarray(a) / b)


def multiply(a, b, name=None):
    return T(np.float32(a) * np.asarray(b)) if np.isscalar(a) else T(np.asarray(a) * np.asarray(b))


def sqrt(x):
    return T(np.sqrt(np.asarray(x)))


def tanh(x):
    return T(np.tanh(np.asarray(x)))


def sigmoid(x):
    a = np.asarray(x)
    return T((1.0 / (1.0 + np.exp(-a))).astype(a.dtype))


def reduce_sum(x, axis=None):
    return T(np.sum(np.asarray(x), axis=axis))


class _NN:
    @staticmethod
    def relu(x, name=None):
        return T(np.maximum(np.asarray(x), 0))

    @staticmethod
    def softmax(x, axis=-1):
        a = np.asarray(x)
        e = np.exp(a - a.max(axis=axis, keepdims=True))
        return T(e / e.sum(axis=axis, keepdims=True))

    @staticmethod
    def moments(x, axes, keep_dims=False):
        a = np.asarray(x)
        m = a.mean(axis=tuple(axes), keepdims=True)
        v = ((a - m) ** 2).mean(axis=tuple(axes), keepdims=True)
        return T(m), T(v)


nn = _NN()


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def extract_image_patches(images, ksizes, strides, rates, padding):
    a = np.asarray(images)
    b, h, w, c = a.shape
    kh, kw = ksizes[1], ksizes[2]
    assert padding == "SAME" and strides == [1, 1, 1, 1] and rates == [1, 1, 1, 1]
    pt, pb = same_pad(h, kh, 1)
    pl, pr = same_pad(w, kw, 1)
    p = np.pad(a, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    cols = [p[:, ky:ky + h, kx:kx + w, :] for ky in np.arange(kh) for kx in np.arange(kw)]
    return T(np.concatenate(cols, axis=-1))  # depth order (ky, kx, c)


class _Image:
    @staticmethod
    def resize_bilinear(x, size, align_corners=False, name=None):
        assert align_corners
        a = np.asarray(x)
        b, h, w, c = a.shape
        oh, ow = [int(v) for v in np.asarray(size).tolist()]
        sy = np.float32((h - 1) / (oh - 1)) if oh > 1 else np.float32(0)
        sx = np.float32((w - 1) / (ow - 1)) if ow > 1 else np.float32(0)
        fy = np.arange(oh, dtype=np.float32) * sy
        fx = np.arange(ow, dtype=np.float32) * sx
        y0 = np.floor(fy).astype(np.int64); x0 = np.floor(fx).astype(np.int64)
        y1 = np.minimum(y0 + 1, h - 1); x1 = np.minimum(x0 + 1, w - 1)
        ly = (fy - y0.astype(np.float32))[None, :, None, None]
        lx = (fx - x0.astype(np.float32))[None, None, :, None]
        tl, tr = a[:, y0][:, :, x0], a[:, y0][:, :, x1]
        bl, br = a[:, y1][:, :, x0], a[:, y1][:, :, x1]
        top = tl + (tr - tl) * lx
        bot = bl + (br - bl) * lx
        return T((top + (bot - top) * ly).astype(np.float32))


image = _Image()


def placeholder(dtype, shape, name=None):
    raise RuntimeError("the numpy shim is eager: call network_graph on arrays")
