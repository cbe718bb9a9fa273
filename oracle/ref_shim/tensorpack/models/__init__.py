"""tensorpack layers used by the reference (Conv2D, AvgPooling, BatchNorm, InstanceNorm, Dropout) with
their documented defaults: Conv2D padding 'same', use_bias True, NHWC, kernel 'W' HWIO + bias 'b';
AvgPooling padding 'valid'; BatchNorm/InstanceNorm epsilon 1e-5, biased variance; BN inference uses
'mean/EMA' and 'variance/EMA'."""
import numpy as np
import tensorflow as tf

__all__ = ["Conv2D", "AvgPooling", "BatchNorm", "InstanceNorm", "Dropout", "layer_register"]


def _pair(k):
    return (k, k) if isinstance(k, int) else tuple(k)


def Conv2D(name, x, filters, kernel_size, strides=1, padding="same", activation=None, use_bias=True):
    a = np.asarray(x, dtype=np.float32)
    kh, kw = _pair(kernel_size)
    s = strides if isinstance(strides, int) else strides[0]
    W = tf.VARIABLES[tf.current_scope(name) + "/W"]
    assert W.shape == (kh, kw, a.shape[3], filters), (name, W.shape, (kh, kw, a.shape[3], filters))
    b, h, w, c = a.shape
    assert padding.lower() == "same"
    pt, pb = tf.same_pad(h, kh, s)
    pl, pr = tf.same_pad(w, kw, s)
    p = np.pad(a, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    oh, ow = -(-h // s), -(-w // s)
    cols = [p[:, ky:ky + (oh - 1) * s + 1:s, kx:kx + (ow - 1) * s + 1:s, :] for ky in range(kh) for kx in range(kw)]
    col = np.concatenate(cols, axis=-1).reshape(b * oh * ow, kh * kw * c)
    y = (col @ W.reshape(kh * kw * c, filters)).reshape(b, oh, ow, filters)
    if use_bias:
        y = y + tf.VARIABLES[tf.current_scope(name) + "/b"]
    y = tf.T(y.astype(np.float32))
    return activation(y) if activation is not None else y


def AvgPooling(name, x, pool_size, strides=None, padding="valid"):
    assert padding == "valid" and pool_size == 2 and strides == 2
    a = np.asarray(x)
    b, h, w, c = a.shape
    oh, ow = h // 2, w // 2
    a = a[:, :oh * 2, :ow * 2, :]
    return tf.T((a[:, 0::2, 0::2] + a[:, 0::2, 1::2] + a[:, 1::2, 0::2] + a[:, 1::2, 1::2]) * np.float32(0.25))


def BatchNorm(name, x, epsilon=1e-5):
    s = tf.current_scope(name)
    g, bt = tf.VARIABLES[s + "/gamma"], tf.VARIABLES[s + "/beta"]
    mu, var = tf.VARIABLES[s + "/mean/EMA"], tf.VARIABLES[s + "/variance/EMA"]
    a = np.asarray(x)
    return tf.T(((a - mu) / np.sqrt(var + np.float32(epsilon)) * g + bt).astype(np.float32))


def InstanceNorm(name, x, epsilon=1e-5, center=True, scale=True):
    assert not center and not scale
    a = np.asarray(x)
    m = a.mean(axis=(1, 2), keepdims=True)
    v = ((a - m) ** 2).mean(axis=(1, 2), keepdims=True)
    return tf.T(((a - m) / np.sqrt(v + np.float32(epsilon))).astype(np.float32))


def Dropout(x, keep_prob=None):
    return x


def layer_register(*a, **k):
    def deco(fn):
        def wrapped(name, *args, **kw):
            return fn(*args, **kw)
        return wrapped
    return deco
