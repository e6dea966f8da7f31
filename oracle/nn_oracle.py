"""CPU oracle for the layer library and models (utils/nn.py, models/dcgan.py, models/densenet.py) -- TEST
INFRASTRUCTURE, NOT PRODUCT.  numpy (fp64 by default) restatement of the TensorFlow ops the reference calls, written
independently of otgan_b200/utils/nn.py (explicit tap loops instead of a conv library call).

PARITY UNPINNED: the reference has no tests and needs TensorFlow 1.x (not installable here); these functions follow the
documented TF-1.x semantics of tf.nn.conv2d('SAME'), tf.nn.l2_normalize, tf.image.resize_nearest_neighbor, tf.split.

Reference lines: get_params utils/nn.py:164-181; apply_pre_activation :190-206; __list_conv2d :234-241; dense :315-325;
conv2d :328-338; adam_updates :50-73; disc_spec/gen_spec models/dcgan.py:7-52, models/densenet.py:7-88.
"""
import numpy as np


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv2d_same(x, W, stride):
    """tf.nn.conv2d(x[NHWC], W[HWIO], [1,s,s,1], 'SAME')"""
    B, H, Wd, C = x.shape
    kh, kw, ci, co = W.shape
    assert ci == C
    oh, pt, pb = same_pad(H, kh, stride)
    ow, pl, pr = same_pad(Wd, kw, stride)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    y = np.zeros((B, oh, ow, co), dtype=x.dtype)
    for a in range(kh):
        for b in range(kw):
            patch = xp[:, a:a + stride * (oh - 1) + 1:stride, b:b + stride * (ow - 1) + 1:stride, :]
            y += np.tensordot(patch, W[a, b], axes=([3], [0]))
    return y


def l2_normalize(v, axes):
    return v / np.sqrt(np.maximum(np.sum(v * v, axis=tuple(axes), keepdims=True), 1e-12))


def weight(V, g):
    """W = l2_normalize(V, all-but-last) * g   (utils/nn.py:176-180)"""
    return l2_normalize(V, range(V.ndim - 1)) * g.reshape([1] * (V.ndim - 1) + [-1])


def pre_activation(xs, kind, axis=3):
    if not isinstance(xs, list):
        xs = [xs]
    if kind is None:
        return np.concatenate(xs, axis)
    if kind == "crelu":
        return np.maximum(np.concatenate([t for x in xs for t in (x, -x)], axis), 0.0)
    if kind == "relu":
        return np.maximum(np.concatenate(xs, axis), 0.0)
    raise ValueError(kind)


def resize_nn(x, oh, ow):
    B, H, W, C = x.shape
    ih = (np.arange(oh) * H) // oh
    iw = (np.arange(ow) * W) // ow
    return x[:, ih][:, :, iw]


class Params:
    """Reads variables by TensorFlow name from a dict {name: array}."""

    def __init__(self, scope, table):
        self.scope, self.table, self.counters = scope, table, {}

    def layer(self, kind):
        k = self.counters.get(kind, 0)
        self.counters[kind] = k + 1
        base = "%s/%s_%d" % (self.scope, kind, k)
        return self.table[base + "/V"], self.table[base + "/g"], self.table[base + "/b"]


def conv_layer(p, x, pre=None, stride=1):
    V, g, b = p.layer("conv2d")
    return conv2d_same(pre_activation(x, pre), weight(V, g), stride) + b


def dense_layer(p, x, pre=None):
    V, g, b = p.layer("dense")
    return pre_activation(x, pre, 1) @ weight(V, g) + b


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def head(x):
    x = np.concatenate([np.maximum(x, 0), np.maximum(-x, 0)], 3)
    x = x.reshape(x.shape[0], -1)
    return x / np.sqrt(np.sum(x * x, axis=1, keepdims=True))


def dcgan_discriminator(x, table):
    p = Params("discriminator", table)
    x = conv_layer(p, x)
    x = conv_layer(p, x, "crelu", 2)
    x = conv_layer(p, x, "crelu", 2)
    x = conv_layer(p, x, "crelu", 2)
    return head(x)


def dcgan_generator(u, table):
    p = Params("generator", table)
    B = u.shape[0]
    x = dense_layer(p, u)
    x, l = np.split(x, 2, 1)
    x = (x * sigmoid(l)).reshape(B, 4, 4, 1024)
    for size in (8, 16, 32):
        x = resize_nn(x, size, size)
        x = conv_layer(p, x)
        x, l = np.split(x, 2, 3)
        x = x * sigmoid(l)
    return np.tanh(conv_layer(p, x))


def densenet_discriminator(x, table, layers_per_block=16, filters_per_layer=16):
    p = Params("discriminator", table)
    x = [conv_layer(p, x)]
    for _ in range(3):
        for _ in range(layers_per_block):
            x.append(conv_layer(p, x, "crelu"))
        x = [conv_layer(p, x, "crelu", 2)]
    return head(x[0])


def densenet_generator(u, table, layers_per_block=16, filters_per_layer=16):
    p = Params("generator", table)
    B = u[0].shape[0]
    x = [dense_layer(p, u[0]).reshape(B, 8, 8, filters_per_layer), u[1]]
    for k in range(3):
        for _ in range(layers_per_block):
            x.append(conv_layer(p, x, "crelu"))
        if k < 2:
            xc = np.concatenate(x, 3)
            xc = resize_nn(xc, xc.shape[1] * 2, xc.shape[2] * 2)
            x = [conv_layer(p, xc, "crelu"), u[k + 2]]
    return np.tanh(conv_layer(p, x, "crelu"))


def adam_step(p, g, v, mg, t, lr, mom1, mom2):
    """utils/nn.py:56-72: one Adam update (epsilon inside the root); returns (p, v, mg)."""
    v_t = mom1 * v + (1.0 - mom1) * g
    v_hat = v_t / (1.0 - mom1 ** t)
    mg_t = mom2 * mg + (1.0 - mom2) * np.square(g)
    mg_hat = mg_t / (1.0 - mom2 ** t)
    return p - lr * v_hat / np.sqrt(mg_hat + 1e-8), v_t, mg_t
