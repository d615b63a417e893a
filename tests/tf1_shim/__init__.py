"""A numpy/torch-CPU backed stand-in for the slice of TensorFlow-1.4 + slim +
fjcommon that the reference's hot-path modules use, so that the UNMODIFIED
reference modules (/root/reference/code/{autoencoder,quantizer,probclass,
ms_ssim,ms_ssim_np,bits,bit_counter,bpp_helpers}.py) can be imported and run in
this container to produce golden vectors (tests/golden/make_golden.py).

TEST INFRASTRUCTURE.  Only used where /root/reference exists (this container).

What this pins: the reference's *graph-level* code -- layer order, scopes,
masks, paddings, slicing, operator-precedence quirks, iteration order --
executes literally.  What it does not pin: the TF kernels themselves, whose
semantics are restated here (independently of oracle/, using torch-CPU
primitives: SAME padding from its definition, conv2d_transpose as the autograd
gradient of the SAME forward conv, fused batch norm inference, conv3d VALID,
softmax / softmax-xent, REFLECT pad).

Execution model: ops run eagerly on numpy values and remember how they were
computed, so ``Session.run(t, feed_dict)`` / ``make_callable`` can re-evaluate
a sub-graph for new placeholder values (needed by probclass.PredictionNetwork).
"""
import contextlib
import functools
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

WEIGHTS = {}          # variable name -> ndarray; filled by the golden script


# ----------------------------------------------------------------------------
# tensors
# ----------------------------------------------------------------------------
class TensorShape(tuple):
    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)

    def is_fully_defined(self):
        return True

    def __getitem__(self, i):
        r = tuple.__getitem__(self, i)
        return TensorShape(r) if isinstance(r, tuple) else r


class DType(object):
    def __init__(self, np_dtype):
        self.np = np.dtype(np_dtype)

    def is_compatible_with(self, other):
        other = other.np if isinstance(other, DType) else np.dtype(other)
        return self.np == other

    @property
    def as_numpy_dtype(self):
        return self.np.type


class T(object):
    """Eager value + provenance (fn, args, kwargs) for re-evaluation.  NOT an
    ndarray subclass: the reference branches on isinstance(x, np.ndarray)
    (probclass.py:272) to tell tensors from numpy inputs."""
    __array_ufunc__ = None          # make ``ndarray <op> T`` defer to T.__r<op>__

    def __init__(self, value):
        self.value = np.asarray(value)
        self._prov = None
        self._placeholder = False

    @property
    def shape(self):
        return TensorShape(self.value.shape)

    @property
    def dtype(self):
        return self.value.dtype

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        assert tuple(shape) == self.value.shape

    def __int__(self):
        return int(self.value)

    def __float__(self):
        return float(self.value)

    def __getitem__(self, idx):
        return _make(lambda a, i: a[i], (self, idx), {})


def _binop(name, fn):
    def f(self, other):
        return _make(fn, (self, other), {})

    def r(self, other):
        return _make(lambda a, b: fn(b, a), (self, other), {})
    setattr(T, '__%s__' % name, f)
    setattr(T, '__r%s__' % name, r)


for _n, _f in (('add', np.add), ('sub', np.subtract), ('mul', np.multiply),
               ('truediv', np.true_divide), ('pow', np.power)):
    _binop(_n, _f)
T.__neg__ = lambda self: _make(np.negative, (self,), {})
T.__isub__ = T.__sub__


def _wrap(v):
    return T(v)


def _plain(v):
    if isinstance(v, T):
        return v.value
    if isinstance(v, (list, tuple)):
        return type(v)(_plain(x) for x in v)
    return v


def _make(raw, args, kwargs):
    val = raw(*_plain(tuple(args)), **{k: _plain(v) for k, v in kwargs.items()})
    t = _wrap(val)
    t._prov = (raw, tuple(args), dict(kwargs))
    return t


def op(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        kwargs.pop('name', None)
        return _make(fn, args, kwargs)
    return wrapper


def _evaluate(t, feeds, memo):
    if isinstance(t, T):
        key = id(t)
        if key in memo:
            return memo[key]
        if t._placeholder:
            for ph, v in feeds:
                if ph is t:
                    memo[key] = np.asarray(v)
                    return memo[key]
            raise ValueError('placeholder not fed')
        if t._prov is None:
            r = t.value
        else:
            raw, args, kwargs = t._prov
            r = np.asarray(raw(*[_evaluate(a, feeds, memo) for a in args],
                               **{k: _evaluate(v, feeds, memo) for k, v in kwargs.items()}))
        memo[key] = r
        return r
    if isinstance(t, (list, tuple)):
        return type(t)(_evaluate(x, feeds, memo) for x in t)
    return t


class Session(object):
    def run(self, fetches, feed_dict=None):
        feeds = list((feed_dict or {}).items())
        memo = {}
        if isinstance(fetches, dict):
            return {k: _evaluate(v, feeds, memo) for k, v in fetches.items()}
        return _evaluate(fetches, feeds, memo)

    def make_callable(self, fetches, feed_list=None):
        def call(*vals):
            return self.run(fetches, dict_items(feed_list, vals))
        return call

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class dict_items(object):       # feed dict keyed by identity (T is unhashable-by-value)
    def __init__(self, keys, vals):
        self._items = list(zip(keys, vals))

    def items(self):
        return self._items


# ----------------------------------------------------------------------------
# scopes, variables
# ----------------------------------------------------------------------------
_scope_stack = []


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None):
    name = name_or_scope if name_or_scope is not None else default_name
    _scope_stack.append(name)
    try:
        yield
    finally:
        _scope_stack.pop()


@contextlib.contextmanager
def name_scope(*a, **k):
    yield


def _full(name):
    return '/'.join([s for s in _scope_stack if s] + [name])


def get_variable(name, shape=None, dtype=None, initializer=None):
    full = _full(name)
    if full not in WEIGHTS:
        raise KeyError('tf1_shim: no value for variable ' + full)
    v = np.asarray(WEIGHTS[full])
    assert shape is None or tuple(shape) == v.shape, (full, shape, v.shape)
    return _wrap(v.copy())


# ----------------------------------------------------------------------------
# slim.arg_scope machinery
# ----------------------------------------------------------------------------
_arg_scopes = [{}]


def add_arg_scope(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        merged = dict(_arg_scopes[-1].get(wrapper, {}))
        merged.update(kwargs)
        return fn(*args, **merged)
    wrapper._orig = fn
    return wrapper


@contextlib.contextmanager
def arg_scope(fns, **kwargs):
    new = {k: dict(v) for k, v in _arg_scopes[-1].items()}
    for f in fns:
        new.setdefault(f, {}).update(kwargs)
    _arg_scopes.append(new)
    try:
        yield
    finally:
        _arg_scopes.pop()


# ----------------------------------------------------------------------------
# TF kernel restatements (torch-CPU primitives; independent of oracle/)
# ----------------------------------------------------------------------------
def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _same_pad(n, k, s):
    out = (n + s - 1) // s
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


def _conv2d_nchw_same(x, w_hwio, stride):
    kh, kw = w_hwio.shape[:2]
    pt, pb = _same_pad(x.shape[2], kh, stride)
    pl, pr = _same_pad(x.shape[3], kw, stride)
    xt = F.pad(_t(x), (pl, pr, pt, pb))
    return F.conv2d(xt, _t(w_hwio.transpose(3, 2, 0, 1)), stride=stride).numpy()


def _conv2d_transpose_nchw_same(x, w, stride):
    """tf.nn.conv2d_transpose == conv2d_backprop_input of the SAME forward conv
    whose input has spatial size stride * size(x); filter [kh,kw,out_c,in_c]."""
    kh, kw, oc, ic = w.shape
    n, _, h, ww = x.shape
    X = torch.zeros((n, oc, h * stride, ww * stride), dtype=_t(x).dtype, requires_grad=True)
    pt, pb = _same_pad(h * stride, kh, stride)
    pl, pr = _same_pad(ww * stride, kw, stride)
    y = F.conv2d(F.pad(X, (pl, pr, pt, pb)), _t(w.transpose(3, 2, 0, 1)), stride=stride)
    (g,) = torch.autograd.grad(y, X, grad_outputs=_t(x))
    return g.numpy()


def _bn_inference_nchw(x, gamma, beta, mean, var, eps):
    sh = (1, -1, 1, 1)
    inv = (1.0 / np.sqrt(var + np.float32(eps))).astype(x.dtype)
    return ((x - mean.reshape(sh)) * inv.reshape(sh)) * gamma.reshape(sh) + beta.reshape(sh)


def _relu(x):
    return np.maximum(x, 0)


def _slim_conv_impl(transpose, inputs, num_outputs, kernel_size, stride=1, padding='SAME',
                    data_format=None, activation_fn=_relu, normalizer_fn=None,
                    normalizer_params=None, weights_regularizer=None, scope=None):
    assert data_format == 'NCHW' and padding == 'SAME' and normalizer_fn is not None
    assert normalizer_params['is_training'] is False and normalizer_params['fused'] and normalizer_params['scale']
    cin = int(inputs.shape[1])
    kh, kw = kernel_size
    with variable_scope(scope):
        wshape = (kh, kw, num_outputs, cin) if transpose else (kh, kw, cin, num_outputs)
        w = get_variable('weights', shape=wshape)
        with variable_scope('BatchNorm'):
            bn = [get_variable(k, shape=(num_outputs,)) for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')]
    eps = normalizer_params['epsilon']

    def raw(x, w_, g, b, m, v):
        y = _conv2d_transpose_nchw_same(x, w_, stride) if transpose else _conv2d_nchw_same(x, w_, stride)
        y = _bn_inference_nchw(y, g, b, m, v, eps)
        return activation_fn(y) if activation_fn is not None else y
    return _make(raw, (inputs, w) + tuple(bn), {})


@add_arg_scope
def slim_conv2d(inputs, num_outputs, kernel_size, **kw):
    return _slim_conv_impl(False, inputs, num_outputs, kernel_size, **kw)


@add_arg_scope
def slim_conv2d_transpose(inputs, num_outputs, kernel_size, **kw):
    return _slim_conv_impl(True, inputs, num_outputs, kernel_size, **kw)


@add_arg_scope
def slim_batch_norm(*a, **k):
    raise NotImplementedError('only used through normalizer_fn')


@op
def nn_conv2d(x, filt, strides, padding):
    assert padding == 'VALID' and list(strides) == [1, 1, 1, 1]
    f = np.asarray(filt).astype(x.dtype)                     # numpy kernel -> tensor dtype
    y = F.conv2d(_t(x.transpose(0, 3, 1, 2)), _t(f.transpose(3, 2, 0, 1)))
    return y.numpy().transpose(0, 2, 3, 1)


@op
def nn_conv3d(x, filt, strides, padding):
    assert padding == 'VALID' and list(strides) == [1, 1, 1, 1, 1]
    y = F.conv3d(_t(x.transpose(0, 4, 1, 2, 3)), _t(filt.transpose(4, 3, 0, 1, 2)))
    return y.numpy().transpose(0, 2, 3, 4, 1)


@op
def nn_softmax(x, dim=-1):
    e = np.exp(x - x.max(axis=dim, keepdims=True))
    return e / e.sum(axis=dim, keepdims=True)


@op
def nn_softmax_xent(logits=None, labels=None):
    sh = logits - logits.max(axis=-1, keepdims=True)
    logp = sh - np.log(np.exp(sh).sum(axis=-1, keepdims=True))
    return -(labels * logp).sum(axis=-1)


@op
def pad(x, pads, mode='CONSTANT', constant_values=0):
    pads = [tuple(int(v) for v in p) for p in pads]
    if mode.upper() == 'REFLECT':
        return np.pad(x, pads, mode='reflect')
    return np.pad(x, pads, mode='constant', constant_values=np.asarray(constant_values, x.dtype))


@op
def one_hot(idx, depth, axis=-1, dtype=np.float32):
    assert axis == -1
    return np.eye(depth, dtype=dtype.np if isinstance(dtype, DType) else dtype)[idx]


def build_modules():
    tf = types.ModuleType('tensorflow')
    tf.float32, tf.int64, tf.uint8, tf.int32 = DType(np.float32), DType(np.int64), DType(np.uint8), DType(np.int32)
    tf.Variable = T
    tf.Tensor = T
    tf.Session = Session
    tf.variable_scope, tf.name_scope, tf.get_variable = variable_scope, name_scope, get_variable
    tf.GraphKeys = types.SimpleNamespace(UPDATE_OPS='update_ops', REGULARIZATION_LOSSES='reg', TRAINABLE_VARIABLES='tv')
    tf.losses = types.SimpleNamespace(add_loss=lambda *a, **k: None, get_regularization_loss=lambda **k: None)
    tf.random_uniform_initializer = lambda **k: None
    tf.zeros_initializer = lambda **k: None
    tf.trainable_variables = lambda scope=None: []
    tf.get_collection = lambda *a, **k: []

    def placeholder(dtype, shape=None, name=None):
        shp = tuple(1 if s is None else int(s) for s in shape)
        t = _wrap(np.zeros(shp, dtype.np))
        t._placeholder = True
        return t
    tf.placeholder = placeholder

    def simple(fn):
        return op(fn)
    tf.expand_dims = simple(lambda x, axis=None, dim=None: np.expand_dims(x, axis if axis is not None else dim))
    tf.concat = simple(lambda vals, axis: np.concatenate(vals, axis=axis))
    tf.transpose = simple(lambda x, perm: np.transpose(x, perm))
    tf.to_float = simple(lambda x: np.asarray(x).astype(np.float32))
    tf.tile = simple(lambda x, m: np.tile(x, m))
    tf.reshape = simple(lambda x, s: np.reshape(x, [int(v) for v in s]))
    tf.reduce_sum = simple(lambda x, axis=None: x.sum(axis=axis, dtype=x.dtype))
    tf.reduce_prod = simple(lambda x, axis=None: np.prod(x, axis=axis, dtype=x.dtype))
    tf.reduce_mean = simple(lambda x, axis=None: x.mean(axis=axis, dtype=x.dtype))
    tf.stop_gradient = simple(lambda x: x)
    tf.identity = simple(lambda x: x)
    tf.stack = simple(lambda vals, axis=0: np.stack(vals, axis=axis))
    tf.shape = simple(lambda x: np.array(x.shape, np.int32))
    tf.gather = simple(lambda params, idx: params[idx])
    tf.constant = lambda v, dtype=None: _wrap(np.asarray(v, dtype.np if dtype is not None else None))
    tf.convert_to_tensor = lambda v, dtype=None: _wrap(np.asarray(v).astype(dtype.np if dtype is not None else np.asarray(v).dtype))
    tf.squeeze = simple(lambda x: np.squeeze(x))
    tf.square = simple(lambda x: np.square(x))
    tf.abs = simple(lambda x: np.abs(x))
    tf.range = lambda n, dtype=None: _wrap(np.arange(n, dtype=dtype.np if dtype is not None else np.int32))
    tf.minimum = simple(lambda a, b: np.minimum(a, np.asarray(b, a.dtype)))
    tf.maximum = simple(lambda a, b: np.maximum(a, np.asarray(b, a.dtype)))
    tf.clip_by_value = simple(lambda x, lo, hi: np.clip(x, np.asarray(lo, x.dtype), np.asarray(hi, x.dtype)))
    tf.cast = simple(lambda x, dtype: x.astype(dtype.np))
    tf.argmax = simple(lambda x, axis=None: np.argmax(x, axis=axis).astype(np.int64))
    tf.add_n = simple(lambda vals: functools.reduce(lambda a, b: a + b, vals))
    tf.pad, tf.one_hot = pad, one_hot

    def py_func(fn, inp, Tout, stateful=True, name=None):
        return _make(lambda *a: np.asarray(fn(*a)), tuple(inp), {})
    tf.py_func = py_func
    tf.nn = types.SimpleNamespace(
        conv2d=nn_conv2d, conv3d=nn_conv3d, softmax=nn_softmax,
        softmax_cross_entropy_with_logits=nn_softmax_xent,
        relu=op(lambda x: np.maximum(x, 0)), sigmoid=op(lambda x: (1 / (1 + np.exp(-x))).astype(x.dtype)),
        bias_add=op(lambda x, b: x + b), l2_loss=op(lambda x: (x * x).sum() / 2))

    contrib = types.ModuleType('tensorflow.contrib')
    slim = types.ModuleType('tensorflow.contrib.slim')
    slim.conv2d, slim.conv2d_transpose, slim.batch_norm = slim_conv2d, slim_conv2d_transpose, slim_batch_norm
    slim.arg_scope, slim.add_arg_scope = arg_scope, add_arg_scope
    slim.l2_regularizer = lambda s: None
    layers = types.ModuleType('tensorflow.contrib.layers')
    layers.xavier_initializer = lambda **k: None
    contrib.slim, contrib.layers = slim, layers
    tf.contrib = contrib

    fj = types.ModuleType('fjcommon')
    th = types.ModuleType('fjcommon.tf_helpers')
    th.transpose_NHWC_to_NCHW = lambda x: tf.transpose(x, (0, 3, 1, 2))
    th.transpose_NCHW_to_NHWC = lambda x: tf.transpose(x, (0, 2, 3, 1))

    def assert_ndims(t, n):
        assert t.shape.ndims == n
    th.assert_ndims = assert_ndims

    def assert_equal_shape(a, b):
        assert tuple(a.shape) == tuple(b.shape), (a.shape, b.shape)
    th.assert_equal_shape = assert_equal_shape
    fe = types.ModuleType('fjcommon.functools_ext')
    fe.identity = lambda x: x
    fe.compose = lambda *fs: functools.reduce(lambda f, g: lambda *a, **k: f(g(*a, **k)), fs)
    tm = types.ModuleType('fjcommon.timer')
    tm.execute = lambda *a, **k: contextlib.nullcontext()
    no = types.ModuleType('fjcommon.no_op')

    class NoOp(object):
        def __call__(self, *a, **k):
            return None
    no.NoOp = NoOp
    fj.tf_helpers, fj.functools_ext, fj.timer, fj.no_op = th, fe, tm, no
    return {'tensorflow': tf, 'tensorflow.contrib': contrib, 'tensorflow.contrib.slim': slim,
            'tensorflow.contrib.layers': layers, 'fjcommon': fj, 'fjcommon.tf_helpers': th,
            'fjcommon.functools_ext': fe, 'fjcommon.timer': tm, 'fjcommon.no_op': no}


def install(reference_code_dir='/root/reference/code'):
    """Put the shim modules in sys.modules and the reference on sys.path."""
    mods = build_modules()
    sys.modules.update(mods)
    if 'scipy.ndimage.filters' not in sys.modules:   # removed from recent scipy; ms_ssim_np.py:22 imports it
        import scipy.ndimage
        m = types.ModuleType('scipy.ndimage.filters')
        m.convolve = scipy.ndimage.convolve
        sys.modules['scipy.ndimage.filters'] = m
    if reference_code_dir not in sys.path:
        sys.path.insert(0, reference_code_dir)
    return mods['tensorflow']
