"""A minimal numpy emulation of the TensorFlow-1.x ops that the REFERENCE's own hot-path files call -- TEST INFRASTRUCTURE.

Purpose: TensorFlow 1.x cannot be installed here (python 3.12, no wheel, no network), so the reference cannot run as
shipped.  What CAN run is the reference's own Python: `utils/matching.py`, `toy_example/matching_cpu.py`, `utils/nn.py`
and `models/*.py` are pure graph-construction code over ~35 TensorFlow primitives.  With this module installed as
`sys.modules['tensorflow']` those files are imported UNMODIFIED from /root/reference and executed eagerly on numpy
arrays (float64 by default): block order, which operand is transposed, the zips / concatenations / regrouping, the
variable creation order, layer sequence, padding and resize calls are then the reference's own code, not a restatement.
Only the primitives below are ours; each states the TensorFlow-1.x semantic it follows.
`make_reference_golden.py` uses it to write tests/golden/ref_*.npz, against which the oracle (CPU) and the CUDA path (GPU)
are checked.

Primitive semantics (TensorFlow 1.x, graph mode, evaluated eagerly here):
  tf.matmul(a, b, transpose_a, transpose_b)        plain matrix product of the (optionally transposed) operands
  tf.concat(values, axis)                          a single (non-list) tensor is wrapped into a list first, i.e. identity
  tf.reduce_logsumexp(x, axis, keep_dims)          log(sum(exp(x - max))) + max, max over the reduced axis (math_ops.py)
  tf.nn.softmax(x)                                 exp(x - max) / sum over the LAST axis
  tf.nn.softmax_cross_entropy_with_logits          -sum(labels * log_softmax(logits), last axis)
  tf.nn.l2_normalize(x, dim)                       x * rsqrt(max(sum(x^2, dim), 1e-12))
  tf.nn.conv2d(x, W, [1,s,s,1], 'SAME')            NHWC x HWIO cross-correlation, out = ceil(n/s),
                                                   pad_total = max((out-1)*s + k - n, 0), before = total // 2
  tf.image.resize_nearest_neighbor(x, [h, w])      align_corners=False: src = floor(dst * in / out)
  tf.nn.moments(x, axes)                           mean, biased variance
  tf.get_variable / variable_scope / make_template name-scoped variable store; a template creates its variables on the
                                                   first call and re-uses them afterwards
  Variable.assign(v)                               returns the assigned VALUE but does not change the stored variable: in
                                                   the reference the init assigns are graph ops that are never fetched
                                                   (train.py builds them under init=True and never runs them)
  tf.random_uniform / random_normal_initializer    drawn from numpy RandomState streams seeded through `set_seed`
"""
import contextlib
import sys
import types

import numpy as np

DTYPE = np.float64          # evaluation precision of the emulated graph (float64: parity target; float32: noise floor)
float32 = "float32"


class _Shape(list):
    def as_list(self):
        return list(self)


class T(np.ndarray):
    """ndarray with the two Tensor methods the reference calls."""

    def get_shape(self):
        return _Shape(int(s) for s in self.shape)

    def set_shape(self, shape):
        assert [int(s) for s in shape] == list(self.shape), (shape, self.shape)

    def assign(self, value):          # see module docstring
        return _t(value)


def _t(x):
    return np.asarray(x, dtype=DTYPE).view(T)


# ------------------------------------------------------------------------------------------------ state
class _State:
    def __init__(self):
        self.reset(0)

    def reset(self, seed):
        self.vars = {}                # full name -> T
        self.order = []               # creation order (== tf.trainable_variables())
        self.scope = []
        self.rng = np.random.RandomState(seed)
        self.overrides = {}           # name -> array installed by the golden generator (seeded weights)


_S = _State()


def set_seed(seed):
    _S.reset(seed)


def variables():
    return [(n, _S.vars[n]) for n in _S.order]


def set_variable(name, value):
    assert name in _S.vars and tuple(_S.vars[name].shape) == tuple(np.shape(value)), name
    _S.vars[name] = _t(value)


# ------------------------------------------------------------------------------------------------ ops
def concat(values, axis=0, name=None):
    if not isinstance(values, (list, tuple)):      # array_ops.concat (TF 1.x): a single tensor is wrapped, i.e. returned as is
        values = [values]
    return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))


def split(value, num_or_size_splits, axis=0):
    return [_t(v) for v in np.split(np.asarray(value), num_or_size_splits, axis=axis)]


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = np.asarray(a), np.asarray(b)
    return _t((a.T if transpose_a else a) @ (b.T if transpose_b else b))


def reduce_logsumexp(x, axis=None, keep_dims=False):
    x = np.asarray(x)
    m = np.max(x, axis=axis, keepdims=True)
    r = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
    return _t(r if keep_dims else np.squeeze(r, axis=axis))


def reduce_sum(x, axis=None, keep_dims=False):
    return _t(np.sum(np.asarray(x), axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keep_dims))


def reduce_mean(x, axis=None, keep_dims=False):
    return _t(np.mean(np.asarray(x), axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keep_dims))


def square(x):
    return _t(np.square(np.asarray(x)))


def sqrt(x):
    return _t(np.sqrt(np.asarray(x)))


def reshape(x, shape):
    return _t(np.reshape(np.asarray(x), [int(s) for s in shape]))


def eye(n):
    return _t(np.eye(int(n)))


def zeros(shape, dtype=None, name=None):
    return _t(np.zeros(shape))


def stop_gradient(x):
    return x


@contextlib.contextmanager
def device(name):
    yield


@contextlib.contextmanager
def control_dependencies(deps):
    yield


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _S.scope.append(name)
    try:
        yield
    finally:
        _S.scope.pop()


def random_normal_initializer(mean=0.0, stddev=1.0):
    return lambda shape: _S.rng.normal(mean, stddev, size=shape)


def ones_initializer():
    return lambda shape: np.ones(shape)


def zeros_initializer():
    return lambda shape: np.zeros(shape)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    full = "/".join(_S.scope + [name])
    if full not in _S.vars:
        if isinstance(shape, (int, np.integer)):
            shape = [int(shape)]
        shape = [int(s) for s in shape]
        _S.vars[full] = _t(initializer(shape))
        _S.order.append(full)
    return _S.vars[full]


def make_template(name, func):
    def template(*args, **kwargs):
        saved = _S.scope
        _S.scope = [name]
        try:
            return func(*args, **kwargs)
        finally:
            _S.scope = saved
    template.__name__ = name
    return template


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=None):
    return _t(_S.rng.uniform(minval, maxval, size=[int(s) for s in shape]))


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _conv2d(x, W, strides, padding, data_format="NHWC"):
    import torch
    import torch.nn.functional as F
    assert padding == "SAME" and strides[0] == 1 and strides[3] == 1
    x, W = np.asarray(x, dtype=DTYPE), np.asarray(W, dtype=DTYPE)
    sh, sw = int(strides[1]), int(strides[2])
    pt, pb = _same_pad(x.shape[1], W.shape[0], sh)
    pl, pr = _same_pad(x.shape[2], W.shape[1], sw)
    xt = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))
    wt = torch.from_numpy(np.ascontiguousarray(W.transpose(3, 2, 0, 1)))
    y = F.conv2d(F.pad(xt, (pl, pr, pt, pb)), wt, stride=(sh, sw))
    return _t(y.numpy().transpose(0, 2, 3, 1))


def _resize_nearest_neighbor(x, size, align_corners=False):
    x = np.asarray(x)
    oh, ow = int(size[0]), int(size[1])
    ih = (np.arange(oh) * x.shape[1]) // oh
    iw = (np.arange(ow) * x.shape[2]) // ow
    return _t(x[:, ih][:, :, iw])


def _softmax(x, dim=-1):
    x = np.asarray(x)
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return _t(e / np.sum(e, axis=-1, keepdims=True))


def _softmax_xent(labels=None, logits=None, **kw):
    l = np.asarray(logits)
    z = l - np.max(l, axis=-1, keepdims=True)
    logsm = z - np.log(np.sum(np.exp(z), axis=-1, keepdims=True))
    return _t(-np.sum(np.asarray(labels) * logsm, axis=-1))


def _l2_normalize(x, dim, epsilon=1e-12):
    x = np.asarray(x)
    ss = np.sum(np.square(x), axis=tuple(int(d) for d in dim) if isinstance(dim, (list, tuple)) else dim, keepdims=True)
    return _t(x / np.sqrt(np.maximum(ss, epsilon)))


def _moments(x, axes):
    x = np.asarray(x)
    ax = tuple(int(a) for a in axes)
    return _t(np.mean(x, axis=ax)), _t(np.var(x, axis=ax))


def _elu(x):
    x = np.asarray(x)
    return _t(np.where(x > 0, x, np.expm1(np.minimum(x, 0))))


# ------------------------------------------------------------------------------------------------ arg_scope (tf.contrib)
_arg_stack = []


def add_arg_scope(fn):
    def wrapped(*args, **kwargs):
        merged = {}
        for names, kw in _arg_stack:
            if fn.__name__ in names:
                merged.update(kw)
        merged.update(kwargs)
        return fn(*args, **merged)
    wrapped.__name__ = fn.__name__
    return wrapped


@contextlib.contextmanager
def arg_scope(funcs, **kwargs):
    _arg_stack.append(({f.__name__ for f in funcs}, kwargs))
    try:
        yield
    finally:
        _arg_stack.pop()


# ------------------------------------------------------------------------------------------------ module assembly
def install():
    """Registers this emulation as `tensorflow` (+ the two sub-module paths the reference imports).  Returns the module."""
    tf = types.ModuleType("tensorflow")
    me = sys.modules[__name__]
    for k in ("concat", "split", "matmul", "reduce_logsumexp", "reduce_sum", "reduce_mean", "square", "sqrt", "reshape", "eye",
              "zeros", "stop_gradient", "device", "control_dependencies", "variable_scope", "random_normal_initializer",
              "ones_initializer", "zeros_initializer", "get_variable", "make_template", "random_uniform", "float32"):
        setattr(tf, k, getattr(me, k))
    nn = types.ModuleType("tensorflow.nn")
    nn.softmax = _softmax
    nn.softmax_cross_entropy_with_logits = _softmax_xent
    nn.l2_normalize = _l2_normalize
    nn.moments = _moments
    nn.conv2d = _conv2d
    nn.bias_add = lambda x, b: _t(np.asarray(x) + np.asarray(b))
    nn.relu = lambda x: _t(np.maximum(np.asarray(x), 0))
    nn.elu = _elu
    nn.sigmoid = lambda x: _t(1.0 / (1.0 + np.exp(-np.asarray(x))))
    nn.tanh = lambda x: _t(np.tanh(np.asarray(x)))
    tf.nn = nn
    image = types.ModuleType("tensorflow.image")
    image.resize_nearest_neighbor = _resize_nearest_neighbor
    tf.image = image
    mods = {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.image": image}
    for path in ("tensorflow.contrib", "tensorflow.contrib.framework", "tensorflow.contrib.framework.python",
                 "tensorflow.contrib.framework.python.ops", "tensorflow.python", "tensorflow.python.framework",
                 "tensorflow.python.framework.function"):
        mods[path] = types.ModuleType(path)
    mods["tensorflow.contrib.framework.python.ops"].arg_scope = arg_scope
    mods["tensorflow.contrib.framework.python.ops"].add_arg_scope = add_arg_scope
    mods["tensorflow.python.framework"].function = mods["tensorflow.python.framework.function"]
    sys.modules.update(mods)
    return tf
