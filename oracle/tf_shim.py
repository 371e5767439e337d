"""A NumPy stand-in for the TensorFlow 1.x API surface that the reference's graph code uses -- TEST INFRASTRUCTURE ONLY.

TensorFlow 1.x is not installable here, so the reference's `deepsignal/model.py` and `deepsignal/layers.py` cannot be run
as they are.  This module lets them run UNMODIFIED: `install()` puts a fake `tensorflow` package into `sys.modules`
whose ops execute eagerly on float64 NumPy arrays (placeholders take their values from `feed(...)`, variables from a
checkpoint dict keyed by TF variable names).  Constructing the reference's `Model(...)` then evaluates the reference's
own graph wiring -- layer order, scopes and variable names, paddings, channel counts, concat order, which timestep of the
BiLSTM is used -- and `model.activation_logits` holds the result.  `tests/golden/make_model_golden.py` uses it to
produce the golden vectors that pin `oracle/model_oracle.py`; what remains restated (here, once, primitive by primitive)
is the documented semantics of the TF ops themselves:

  * `tf.layers.conv2d` / pooling with padding="SAME": out = ceil(in / stride), pad_total = max((out-1)*stride + k - in, 0),
    pad_before = pad_total // 2; max pooling ignores the padding, average pooling divides by the number of valid elements
  * `tf.contrib.layers.batch_norm` (is_training=False): (x - moving_mean) / sqrt(moving_variance + 0.001) * gamma + beta
  * `tf.contrib.rnn.LSTMCell` (defaults): gates i, j, f, o = split(concat([x, h]) @ kernel + bias, 4);
    c = sigmoid(f + 1.0) * c_prev + sigmoid(i) * tanh(j);  h = sigmoid(o) * tanh(c)
  * `tf.nn.bidirectional_dynamic_rnn`: the backward direction runs on the time-reversed sequence and its outputs are
    reversed back; variables live under <scope>/{fw,bw}/multi_rnn_cell/cell_<i>/lstm_cell/{kernel,bias}
  * `tf.layers.dense` without a name: layers are called dense, dense_1, ... in creation order

Only `tests/`, `tests/golden/make_model_golden.py` and nothing in the product may import this module.
"""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as np

F = np.float64


class Tensor(np.ndarray):
    """ndarray with the two TF tensor methods the reference calls."""

    def get_shape(self):
        shape = tuple(self.shape)

        class _S:
            def as_list(self_inner):
                return list(shape)
        return _S()


def T(a, dtype=None):
    return np.asarray(a, dtype=dtype).view(Tensor)


class _State:
    def __init__(self):
        self.reset({}, [])

    def reset(self, ckpt, feeds):
        self.ckpt = ckpt
        self.feeds = list(feeds)
        self.scope = []            # variable-scope name stack
        self.used = []             # variable names in creation order
        self.dense_count = 0
        self.arg_defaults = []


S = _State()


def begin(ckpt, feeds):
    """ckpt: {tf variable name: array}; feeds: placeholder values in the order the graph code creates its placeholders."""
    S.reset(ckpt, feeds)


def variables_used():
    return list(S.used)


# ---------------------------------------------------------------- scopes / variables
class _VarScope:
    def __init__(self, full):
        self.name = full


@contextlib.contextmanager
def variable_scope(name_or_scope, *a, **k):
    saved = list(S.scope)
    if isinstance(name_or_scope, _VarScope):
        S.scope = [p for p in name_or_scope.name.split("/") if p]
    else:
        S.scope = saved + [str(name_or_scope)]
    try:
        yield _VarScope("/".join(S.scope))
    finally:
        S.scope = saved


@contextlib.contextmanager
def name_scope(*a, **k):          # op names only: variable names ignore it
    yield None


def _var(name, shape=None):
    full = "/".join(S.scope + [name])
    if full not in S.ckpt:
        if full.endswith("global_step"):
            return T(np.zeros((), np.int32))
        raise KeyError("the graph code asks for variable %r, which the checkpoint does not have" % full)
    v = np.asarray(S.ckpt[full])
    if shape is not None and tuple(int(x) for x in shape) != tuple(v.shape):
        raise ValueError("variable %r: graph shape %r, checkpoint shape %r" % (full, tuple(shape), tuple(v.shape)))
    S.used.append(full)
    return T(v, F)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **k):
    return _var(name, shape)


# ---------------------------------------------------------------- basic ops
def placeholder(dtype, shape=None, name=None):
    v = S.feeds.pop(0)
    a = np.asarray(v)
    if shape is not None:
        want = [None if d is None else int(d) for d in shape]
        if len(want) != a.ndim or any(w is not None and w != g for w, g in zip(want, a.shape)):
            raise ValueError("placeholder %r fed with shape %r" % (shape, a.shape))
    if dtype is float32:
        a = a.astype(F)
    elif dtype is int32:
        a = a.astype(np.int64)
    return T(a)


class _DType:
    def __init__(self, name):
        self.name = name


float32, int32, bool_ = _DType("float32"), _DType("int32"), _DType("bool")


def _np_dtype(dt):
    return {float32: F, int32: np.int64, bool_: np.bool_}.get(dt, F)


def reshape(x, shape, name=None):
    return T(np.reshape(np.asarray(x), [int(s) for s in shape]))


def concat(values, axis, name=None):
    return T(np.concatenate([np.asarray(v) for v in values], axis=axis))


def squeeze(x, axis=None, name=None):
    return T(np.squeeze(np.asarray(x), axis=axis))


def shape(x, name=None):
    return T(np.asarray(np.asarray(x).shape, np.int64))


def slice_(x, begin, size, name=None):
    x = np.asarray(x)
    idx = tuple(slice(int(b), x.shape[i] if int(s) == -1 else int(b) + int(s)) for i, (b, s) in enumerate(zip(begin, size)))
    return T(x[idx])


def ones_like(x, dtype=None, name=None):
    return T(np.ones_like(np.asarray(x), dtype=_np_dtype(dtype) if dtype is not None else None))


def zeros_like(x, dtype=None, name=None):
    return T(np.zeros_like(np.asarray(x), dtype=_np_dtype(dtype) if dtype is not None else None))


def cast(x, dtype, name=None):
    return T(np.asarray(x).astype(_np_dtype(dtype)))


def one_hot(x, depth, **k):
    return T(np.eye(int(depth), dtype=F)[np.asarray(x, np.int64)])


def argmax(x, axis=None, **k):
    return T(np.argmax(np.asarray(x), axis=axis))


def where(c, a, b, name=None):
    return T(np.where(np.asarray(c), np.asarray(a), np.asarray(b)))


def greater(a, b, name=None):
    return T(np.asarray(a) > np.asarray(b))


def reduce_mean(x, axis=None, **k):
    return T(np.mean(np.asarray(x), axis=axis))


def add(a, b, name=None):
    return T(np.asarray(a) + np.asarray(b))


def cond(pred, true_fn, false_fn, **k):
    return true_fn() if bool(np.asarray(pred)) else false_fn()


def zeros_initializer(*a, **k):
    return None


def truncated_normal_initializer(*a, **k):
    return None


# ---------------------------------------------------------------- nn
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def relu(x, name=None):
    return T(np.maximum(np.asarray(x), 0.0))


def sigmoid(x, name=None):
    return T(_sigmoid(np.asarray(x, F)))


def dropout(x, keep_prob=None, **k):
    if float(np.asarray(keep_prob)) != 1.0:
        raise NotImplementedError("inference only: keep_prob must be fed as 1.0")
    return x


def embedding_lookup(params, ids, **k):
    return T(np.asarray(params)[np.asarray(ids, np.int64)])


def weighted_cross_entropy_with_logits(targets=None, logits=None, pos_weight=1.0, **k):
    z, x = np.asarray(targets, F), np.asarray(logits, F)
    return T(-(pos_weight * z * np.log(_sigmoid(x)) + (1 - z) * np.log(1 - _sigmoid(x))))


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def _windows(x, kh, kw, stride, fill):
    """x [N,H,W,C] -> (windows [N,Ho,Wo,kh,kw,C], valid mask [Ho,Wo,kh,kw]) with SAME padding."""
    sh, sw = (stride, stride) if np.isscalar(stride) else stride
    n, h, w, c = x.shape
    ho, pt, pb = _same_pad(h, kh, sh)
    wo, pl, pr = _same_pad(w, kw, sw)
    xp = np.full((n, h + pt + pb, w + pl + pr, c), fill, F)
    xp[:, pt:pt + h, pl:pl + w, :] = x
    mp = np.zeros((h + pt + pb, w + pl + pr), bool)
    mp[pt:pt + h, pl:pl + w] = True
    win = np.empty((n, ho, wo, kh, kw, c), F)
    val = np.empty((ho, wo, kh, kw), bool)
    for a in range(kh):
        for b in range(kw):
            win[:, :, :, a, b, :] = xp[:, a:a + (ho - 1) * sh + 1:sh, b:b + (wo - 1) * sw + 1:sw, :]
            val[:, :, a, b] = mp[a:a + (ho - 1) * sh + 1:sh, b:b + (wo - 1) * sw + 1:sw]
    return win, val


def conv2d(inputs, filters, kernel_size, strides=1, padding="SAME", use_bias=True, name=None, **k):
    if padding != "SAME" or use_bias:
        raise NotImplementedError("the reference only uses SAME padding without bias")
    x = np.asarray(inputs, F)
    kh, kw = kernel_size
    with variable_scope(name):
        ker = np.asarray(_var("kernel", (kh, kw, x.shape[-1], filters)))
    win, _ = _windows(x, kh, kw, strides, 0.0)
    return T(np.einsum("nhwabc,abcf->nhwf", win, ker))


def max_pooling2d(inputs, pool_size, strides, padding="SAME", name=None, **k):
    win, _ = _windows(np.asarray(inputs, F), pool_size[0], pool_size[1], strides, -np.inf)
    return T(win.max(axis=(3, 4)))


def average_pooling2d(inputs, pool_size, strides, padding="SAME", name=None, **k):
    win, val = _windows(np.asarray(inputs, F), pool_size[0], pool_size[1], strides, 0.0)
    cnt = val.sum(axis=(2, 3)).astype(F)
    return T(win.sum(axis=(3, 4)) / cnt[None, :, :, None])


def dense(inputs, units, use_bias=True, name=None, **k):
    if use_bias:
        raise NotImplementedError
    if name is None:
        name = "dense" if S.dense_count == 0 else "dense_%d" % S.dense_count
        S.dense_count += 1
    x = np.asarray(inputs, F)
    with variable_scope(name):
        ker = np.asarray(_var("kernel", (x.shape[-1], units)))
    return T(x @ ker)


# ---------------------------------------------------------------- contrib: arg_scope / batch_norm
@contextlib.contextmanager
def arg_scope(funcs, **kwargs):
    S.arg_defaults.append(kwargs)
    try:
        yield
    finally:
        S.arg_defaults.pop()


def batch_norm(inputs=None, is_training=True, reuse=None, scope=None, **kwargs):
    kw = {}
    for d in S.arg_defaults:
        kw.update(d)
    kw.update(kwargs)
    scope = scope if scope is not None else kw.get("scope")
    if bool(np.asarray(is_training)):
        raise NotImplementedError("inference only: `training` must be fed as False")
    if not kw.get("center", True) or not kw.get("scale", False):
        raise NotImplementedError("the reference uses center=True, scale=True")
    x = np.asarray(inputs, F)
    c = x.shape[-1]
    with variable_scope(scope):
        beta, gamma = np.asarray(_var("beta", (c,))), np.asarray(_var("gamma", (c,)))
        mean, var = np.asarray(_var("moving_mean", (c,))), np.asarray(_var("moving_variance", (c,)))
    eps = kw.get("epsilon", 0.001)
    return T((x - mean) / np.sqrt(var + eps) * gamma + beta)


# ---------------------------------------------------------------- contrib.rnn
class LSTMCell:
    def __init__(self, num_units, **k):
        if k:
            raise NotImplementedError("the reference uses LSTMCell defaults")
        self.n = int(num_units)

    def variables(self, in_dim):
        with variable_scope("lstm_cell"):
            return np.asarray(_var("kernel", (in_dim + self.n, 4 * self.n))), np.asarray(_var("bias", (4 * self.n,)))

    def step(self, x, c, h, kernel, bias):
        z = np.concatenate([x, h], axis=1) @ kernel + bias
        i, j, f, o = np.split(z, 4, axis=1)
        c2 = _sigmoid(f + 1.0) * c + _sigmoid(i) * np.tanh(j)
        return c2, _sigmoid(o) * np.tanh(c2)


class GRUCell:
    def __init__(self, *a, **k):
        raise NotImplementedError("cell='GRU' is not used by the reference's Model")


class DropoutWrapper:
    def __init__(self, cell, output_keep_prob=1.0, **k):
        if float(np.asarray(output_keep_prob)) != 1.0:
            raise NotImplementedError("inference only: keep_prob must be fed as 1.0")
        self.cell = cell


class MultiRNNCell:
    def __init__(self, cells):
        self.cells = [c.cell if isinstance(c, DropoutWrapper) else c for c in cells]


def _run_direction(multi, x):
    """x [N, T, D] -> outputs of the top layer [N, T, H]."""
    n, t, _ = x.shape
    with variable_scope("multi_rnn_cell"):
        params = []
        in_dim = x.shape[2]
        for li, cell in enumerate(multi.cells):
            with variable_scope("cell_%d" % li):
                params.append(cell.variables(in_dim))
            in_dim = cell.n
    cs = [np.zeros((n, c.n), F) for c in multi.cells]
    hs = [np.zeros((n, c.n), F) for c in multi.cells]
    out = np.empty((n, t, multi.cells[-1].n), F)
    for step in range(t):
        inp = x[:, step, :]
        for li, cell in enumerate(multi.cells):
            cs[li], hs[li] = cell.step(inp, cs[li], hs[li], *params[li])
            inp = hs[li]
        out[:, step, :] = inp
    return out


def bidirectional_dynamic_rnn(cell_fw=None, cell_bw=None, inputs=None, sequence_length=None, dtype=None, scope=None, **k):
    x = np.asarray(inputs, F)
    if sequence_length is not None and not np.all(np.asarray(sequence_length) == x.shape[1]):
        raise NotImplementedError("the reference feeds full-length sequences")
    with variable_scope(scope if scope is not None else "bidirectional_rnn"):
        with variable_scope("fw"):
            fw = _run_direction(cell_fw, x)
        with variable_scope("bw"):
            bw = _run_direction(cell_bw, x[:, ::-1, :])[:, ::-1, :]
    return (T(fw), T(bw)), None


class _Adam:
    def __init__(self, *a, **k):
        pass

    def minimize(self, *a, **k):
        return None


# ---------------------------------------------------------------- the fake package
def install():
    """Put the fake `tensorflow` (and tensorflow.contrib.*) into sys.modules; returns the previous entries."""
    names = ["tensorflow", "tensorflow.contrib", "tensorflow.contrib.rnn", "tensorflow.contrib.layers", "tensorflow.contrib.framework"]
    saved = {n: sys.modules.get(n) for n in names}
    tf = types.ModuleType("tensorflow")
    for k, v in dict(float32=float32, int32=int32, bool=bool_, placeholder=placeholder, get_variable=get_variable,
                     variable_scope=variable_scope, name_scope=name_scope, reshape=reshape, concat=concat, squeeze=squeeze,
                     shape=shape, slice=slice_, ones_like=ones_like, zeros_like=zeros_like, cast=cast, one_hot=one_hot,
                     argmax=argmax, where=where, greater=greater, reduce_mean=reduce_mean, add=add, cond=cond,
                     zeros_initializer=zeros_initializer, truncated_normal_initializer=truncated_normal_initializer).items():
        setattr(tf, k, v)
    tf.nn = types.SimpleNamespace(relu=relu, sigmoid=sigmoid, dropout=dropout, embedding_lookup=embedding_lookup,
                                  weighted_cross_entropy_with_logits=weighted_cross_entropy_with_logits,
                                  bidirectional_dynamic_rnn=bidirectional_dynamic_rnn,
                                  rnn_cell=types.SimpleNamespace(MultiRNNCell=MultiRNNCell))
    tf.layers = types.SimpleNamespace(conv2d=conv2d, max_pooling2d=max_pooling2d, average_pooling2d=average_pooling2d, dense=dense)
    tf.train = types.SimpleNamespace(AdamOptimizer=_Adam)
    contrib = types.ModuleType("tensorflow.contrib")
    rnn = types.ModuleType("tensorflow.contrib.rnn")
    rnn.LSTMCell, rnn.GRUCell, rnn.DropoutWrapper = LSTMCell, GRUCell, DropoutWrapper
    lay = types.ModuleType("tensorflow.contrib.layers")
    lay.batch_norm = batch_norm
    fw = types.ModuleType("tensorflow.contrib.framework")
    fw.arg_scope = arg_scope
    contrib.rnn, contrib.layers, contrib.framework = rnn, lay, fw
    tf.contrib = contrib
    for n, m in zip(names, [tf, contrib, rnn, lay, fw]):
        sys.modules[n] = m
    return saved


def uninstall(saved):
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m
