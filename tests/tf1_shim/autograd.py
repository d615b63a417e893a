"""torch-autograd stand-in for the TF-1.4 / slim / fjcommon symbols the reference's TRAINING graph uses
(TEST INFRASTRUCTURE, like tests/tf1_shim/__init__.py).

The numpy shim next door executes the unmodified reference modules in inference mode; this sibling executes them --
autoencoder.py, quantizer.py, probclass.py, ms_ssim.py, bits.py and the get_loss / Distortions definitions of train.py
(code/train.py:303-336,352-431) -- with is_training=True on torch-CPU tensors with autograd, so that
`total_loss.t.backward()` plays the role of tf.gradients (code/train.py:339-349).  What it pins: the reference's OWN
graph construction for the training step (layer order, scopes, batch-norm mode, stop_gradient placement, heatmap
routing of the rate term, regularisation terms), executed on restated TF kernel semantics:
SAME padding, conv2d_transpose = gradient of the SAME conv, fused batch norm with batch statistics (biased variance for
the normalisation, unbiased into the moving average), PadV2 (no gradient to constant_values), REFLECT pad,
softmax cross entropy, maximum / minimum / clip_by_value sub-gradients.

    tf = install(weights, dtype=torch.float64)     # puts 'tensorflow', 'fjcommon', ... into sys.modules
    ... build the graph with the reference's classes ...
    state().vars[name].t.grad                       # d loss / d variable after loss.t.backward()
"""
import contextlib
import functools
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F


class TensorShape(tuple):
    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)

    def is_fully_defined(self):
        return True


class DType(object):
    def __init__(self, name, kind, torch_dtype):
        self.name, self.kind, self.torch = name, kind, torch_dtype

    def is_compatible_with(self, other):
        return isinstance(other, DType) and other.kind == self.kind

    def __repr__(self):
        return 'tf.' + self.name


class _State(object):
    def __init__(self):
        self.fdt = torch.float64
        self.weights = {}
        self.vars = {}              # full name -> AT (leaf, requires_grad)
        self.reg_losses = []        # (name path, AT)
        self.bn_stats = {}          # conv scope -> (batch mean, unbiased batch variance)
        self.var_scopes = []
        self.name_scopes = []
        self.arg_scopes = [{}]


_S = _State()


def state():
    return _S


FLOAT32 = DType('float32', 'f', None)
INT64, INT32, UINT8 = DType('int64', 'i', torch.int64), DType('int32', 'i', torch.int32), DType('uint8', 'i', torch.uint8)


class AT(object):
    """autograd tensor: a torch tensor with the few Tensor attributes the reference touches"""
    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, t, name=None):
        self.t = t
        self.name = name or ''

    @property
    def shape(self):
        return TensorShape(int(s) for s in self.t.shape)

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        assert tuple(shape) == tuple(self.t.shape)

    @property
    def dtype(self):
        if self.t.dtype.is_floating_point:
            return FLOAT32
        return {torch.int64: INT64, torch.int32: INT32, torch.uint8: UINT8}.get(self.t.dtype, INT64)

    def __len__(self):
        return self.t.shape[0]

    def __int__(self):
        return int(self.t)

    def __float__(self):
        return float(self.t)

    def __index__(self):
        return int(self.t)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return AT(self.t[tuple(raw(i) if isinstance(i, AT) else i for i in idx)])


def raw(v):
    """AT / ndarray / python number -> torch tensor (floating values in the shim's float type)"""
    if isinstance(v, AT):
        return v.t
    if isinstance(v, torch.Tensor):
        return v
    a = np.asarray(v)
    if a.dtype.kind == 'f':
        return torch.from_numpy(np.ascontiguousarray(a.astype(np.float64))).to(_S.fdt)
    return torch.from_numpy(np.ascontiguousarray(a))


def _bin(fn):
    def f(self, other):
        a, b = raw(self), raw(other)
        if a.dtype.is_floating_point and not b.dtype.is_floating_point:
            b = b.to(a.dtype)
        return AT(fn(a, b))

    def r(self, other):
        a, b = raw(other), raw(self)
        if b.dtype.is_floating_point and not a.dtype.is_floating_point:
            a = a.to(b.dtype)
        return AT(fn(a, b))
    return f, r


for _n, _f in (('add', torch.add), ('sub', torch.sub), ('mul', torch.mul), ('truediv', torch.true_divide), ('pow', torch.pow)):
    _fwd, _rev = _bin(_f)
    setattr(AT, '__%s__' % _n, _fwd)
    setattr(AT, '__r%s__' % _n, _rev)
AT.__neg__ = lambda self: AT(-self.t)
AT.__isub__ = AT.__sub__            # "sigma11 -= mu11" (code/ms_ssim.py:100-102) must not be in place under autograd
AT.__iadd__ = AT.__add__


def _ints(seq):
    if isinstance(seq, AT):
        return [int(v) for v in seq.t.tolist()]
    return [int(v) for v in seq]


def op(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        kwargs.pop('name', None)
        return fn(*args, **kwargs)
    return wrapper


# ------------------------------------------------------------------ scopes, variables, collections
@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None):
    name = name_or_scope if name_or_scope is not None else default_name
    _S.var_scopes.append(name)
    _S.name_scopes.append(name)
    try:
        yield
    finally:
        _S.var_scopes.pop()
        _S.name_scopes.pop()


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    _S.name_scopes.append(name if name is not None else default_name)
    try:
        yield
    finally:
        _S.name_scopes.pop()


def _var_path(name):
    return '/'.join([s for s in _S.var_scopes if s] + [name])


def _name_path():
    return '/'.join(s for s in _S.name_scopes if s)


def get_variable(name, shape=None, dtype=None, initializer=None):
    full = _var_path(name)
    if full not in _S.vars:
        if full not in _S.weights:
            raise KeyError('tf1_shim.autograd: no value for variable ' + full)
        v = np.asarray(_S.weights[full])
        assert shape is None or tuple(int(s) for s in shape) == v.shape, (full, shape, v.shape)
        _S.vars[full] = AT(raw(v.astype(np.float64)).clone().requires_grad_(True), name=full + ':0')
    return _S.vars[full]


def add_loss(loss, collection=None):
    _S.reg_losses.append((_name_path(), loss))


def get_regularization_loss(scope=None, name=None):
    sel = [l for p, l in _S.reg_losses if scope is None or p.startswith(scope)]
    if not sel:
        return AT(torch.zeros((), dtype=_S.fdt))
    return AT(functools.reduce(torch.add, [l.t for l in sel]))


def trainable_variables(scope=None):
    return [v for k, v in _S.vars.items() if scope is None or k.startswith(scope)]


def add_arg_scope(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        merged = dict(_S.arg_scopes[-1].get(wrapper, {}))
        merged.update(kwargs)
        return fn(*args, **merged)
    wrapper._orig = fn
    return wrapper


@contextlib.contextmanager
def arg_scope(fns, **kwargs):
    new = {k: dict(v) for k, v in _S.arg_scopes[-1].items()}
    for f in fns:
        new.setdefault(f, {}).update(kwargs)
    _S.arg_scopes.append(new)
    try:
        yield
    finally:
        _S.arg_scopes.pop()


# ------------------------------------------------------------------ TF kernel semantics
def _same_pad(n, k, s):
    out = (n + s - 1) // s
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


def _conv2d_nchw_same(x, w_hwio, stride):
    kh, kw = w_hwio.shape[:2]
    pt, pb = _same_pad(x.shape[2], kh, stride)
    pl, pr = _same_pad(x.shape[3], kw, stride)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w_hwio.permute(3, 2, 0, 1), stride=stride)


def _conv2d_transpose_nchw_same(x, w, stride):
    """tf.nn.conv2d_transpose = Conv2DBackpropInput of the SAME forward conv whose input has spatial size stride * size(x);
    filter [kh, kw, out_c, in_c].  The full transposed conv, cropped by that forward conv's leading padding."""
    kh, kw = w.shape[:2]
    full = F.conv_transpose2d(x, w.permute(3, 2, 0, 1), stride=stride)
    H, W = x.shape[2] * stride, x.shape[3] * stride
    pt, pl = _same_pad(H, kh, stride)[0], _same_pad(W, kw, stride)[0]
    return full[:, :, pt:pt + H, pl:pl + W]


def _slim_conv_impl(transpose, inputs, num_outputs, kernel_size, stride=1, padding='SAME', data_format=None,
                    activation_fn='relu', normalizer_fn=None, normalizer_params=None, weights_regularizer=None, scope=None):
    assert data_format == 'NCHW' and padding == 'SAME' and normalizer_fn is not None
    p = normalizer_params
    assert p['fused'] and p['scale'] and p['data_format'] == 'NCHW' and p['decay'] == 0.9
    x = raw(inputs)
    cin = int(x.shape[1])
    kh, kw = kernel_size
    with variable_scope(scope):
        conv_scope = _var_path('')[:-1]
        wshape = (kh, kw, num_outputs, cin) if transpose else (kh, kw, cin, num_outputs)
        w = get_variable('weights', shape=wshape)
        if weights_regularizer is not None:              # slim: '<scope>/kernel/Regularizer/l2_regularizer'
            with name_scope('kernel/Regularizer'):
                add_loss(weights_regularizer(w))
        with variable_scope('BatchNorm'):
            gamma, beta = get_variable('gamma', shape=(num_outputs,)), get_variable('beta', shape=(num_outputs,))
            mm, mv = get_variable('moving_mean', shape=(num_outputs,)), get_variable('moving_variance', shape=(num_outputs,))
    y = _conv2d_transpose_nchw_same(x, w.t, stride) if transpose else _conv2d_nchw_same(x, w.t, stride)
    sh = (1, -1, 1, 1)
    if p['is_training']:            # FusedBatchNorm(is_training=True): batch mean / biased variance
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        n = y.numel() // y.shape[1]
        _S.bn_stats[conv_scope] = (mean.detach().numpy().copy(), (var * n / max(n - 1, 1)).detach().numpy().copy())
    else:
        mean, var = mm.t.detach(), mv.t.detach()
    y = (y - mean.reshape(sh)) * (torch.rsqrt(var + p['epsilon']) * gamma.t).reshape(sh) + beta.t.reshape(sh)
    if activation_fn == 'relu' or activation_fn is relu:
        y = torch.relu(y)
    else:
        assert activation_fn is None, activation_fn
    return AT(y)


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
def relu(x):
    return AT(torch.relu(raw(x)))


@op
def nn_conv2d(x, filt, strides, padding):
    assert padding == 'VALID' and list(strides) == [1, 1, 1, 1]
    xt = raw(x)
    f = raw(np.asarray(filt, np.float64)).to(xt.dtype)          # HWIO
    return AT(F.conv2d(xt.permute(0, 3, 1, 2), f.permute(3, 2, 0, 1)).permute(0, 2, 3, 1))


@op
def nn_conv3d(x, filt, strides, padding):
    assert padding == 'VALID' and list(strides) == [1, 1, 1, 1, 1]
    return AT(F.conv3d(raw(x).permute(0, 4, 1, 2, 3), raw(filt).permute(4, 3, 0, 1, 2)).permute(0, 2, 3, 4, 1))


@op
def nn_softmax(x, dim=-1):
    return AT(torch.softmax(raw(x), dim=dim))


@op
def nn_softmax_xent(logits=None, labels=None):
    return AT(-(raw(labels) * torch.log_softmax(raw(logits), dim=-1)).sum(-1))


@op
def pad(x, pads, mode='CONSTANT', constant_values=0):
    xt = raw(x)
    pads = [tuple(int(v) for v in p) for p in pads]
    flat = []
    for lo, hi in reversed(pads):
        flat += [lo, hi]
    if mode.upper() == 'REFLECT':
        # F.pad(reflect) handles the last 2 (3) dims of a 4-d (5-d) tensor: the reference pads H and W of NHWC
        assert xt.dim() == 4 and pads[0] == (0, 0) and pads[3] == (0, 0)
        (t, b), (l, r) = pads[1], pads[2]
        if t == b == l == r == 0:
            return AT(xt)
        y = F.pad(xt.permute(0, 3, 1, 2), (l, r, t, b), mode='reflect')
        return AT(y.permute(0, 2, 3, 1))
    cv = constant_values
    cv = float(raw(cv).detach()) if isinstance(cv, (AT, torch.Tensor)) else float(cv)     # PadV2: no gradient to constant_values
    return AT(F.pad(xt, flat, mode='constant', value=cv))


@op
def one_hot(idx, depth, axis=-1, dtype=None):
    assert axis == -1
    return AT(F.one_hot(raw(idx).long(), int(depth)).to(_S.fdt))


def _axis(axis):
    if axis is None:
        return None
    return tuple(int(a) for a in axis) if isinstance(axis, (list, tuple)) else int(axis)


def _reduce(fn):
    @op
    def f(x, axis=None, keep_dims=False):
        t = raw(x)
        if axis is None:
            return AT(fn(t))
        return AT(fn(t, dim=_axis(axis), keepdim=keep_dims))
    return f


def _prod(t, dim=None, keepdim=False):
    if dim is None:
        return t.prod()
    for d in sorted((dim,) if isinstance(dim, int) else dim, reverse=True):
        t = t.prod(dim=d, keepdim=keepdim)
    return t


def _cast(x, dtype):
    t = raw(x)
    if dtype.kind == 'f':
        return AT(t.to(_S.fdt))
    return AT(t.to(dtype.torch))          # float -> int truncates toward zero like tf.cast


def build_modules():
    tf = types.ModuleType('tensorflow')
    tf.float32, tf.int64, tf.int32, tf.uint8 = FLOAT32, INT64, INT32, UINT8
    tf.Variable = tf.Tensor = AT
    tf.variable_scope, tf.name_scope, tf.get_variable = variable_scope, name_scope, get_variable
    tf.GraphKeys = types.SimpleNamespace(UPDATE_OPS='update_ops', REGULARIZATION_LOSSES='reg', TRAINABLE_VARIABLES='tv')
    tf.losses = types.SimpleNamespace(add_loss=add_loss, get_regularization_loss=get_regularization_loss)
    tf.random_uniform_initializer = lambda **k: None
    tf.zeros_initializer = lambda **k: None
    tf.trainable_variables = trainable_variables
    tf.get_collection = lambda key, scope=None: trainable_variables(scope)
    tf.summary = types.SimpleNamespace(scalar=lambda *a, **k: None)
    tf.expand_dims = op(lambda x, axis=None, dim=None: AT(raw(x).unsqueeze(axis if axis is not None else dim)))
    tf.concat = op(lambda vals, axis: AT(torch.cat([raw(v) for v in vals], dim=axis)))
    tf.stack = op(lambda vals, axis=0: AT(torch.stack([raw(v) for v in vals], dim=axis)))
    tf.transpose = op(lambda x, perm: AT(raw(x).permute(*[int(p) for p in perm])))
    tf.to_float = op(lambda x: AT(raw(x).to(_S.fdt)))
    tf.tile = op(lambda x, m: AT(raw(x).repeat(*_ints(m))))
    tf.reshape = op(lambda x, s: AT(raw(x).reshape(_ints(s))))
    tf.reduce_sum, tf.reduce_mean = _reduce(torch.sum), _reduce(torch.mean)
    tf.reduce_prod = _reduce(_prod)
    tf.stop_gradient = op(lambda x: AT(raw(x).detach()))
    tf.identity = op(lambda x: AT(raw(x)))
    tf.shape = op(lambda x: AT(torch.tensor(list(raw(x).shape), dtype=torch.int64)))
    tf.gather = op(lambda params, idx: AT(raw(params)[raw(idx).long()]))
    tf.constant = op(lambda v, dtype=None, shape=None: AT(raw(v) if dtype is None else _cast(raw(v), dtype).t))
    tf.convert_to_tensor = op(lambda v, dtype=None: AT(raw(v) if dtype is None else _cast(raw(v), dtype).t))
    tf.squeeze = op(lambda x, axis=None: AT(raw(x).squeeze() if axis is None else raw(x).squeeze(axis)))
    tf.square = op(lambda x: AT(raw(x) * raw(x)))
    tf.abs = op(lambda x: AT(raw(x).abs()))
    tf.range = op(lambda n, dtype=None: AT(torch.arange(int(n), dtype=_S.fdt if (dtype is None or dtype.kind == 'f') else dtype.torch)))
    # sub-gradients: TF sends the gradient of maximum(x, y) to x where x >= y (minimum: x <= y); torch.where reproduces it
    tf.maximum = op(lambda a, b: AT((lambda x, y: torch.where(x >= y, x, y))(*torch.broadcast_tensors(raw(a), raw(b).to(raw(a).dtype)))))
    tf.minimum = op(lambda a, b: AT((lambda x, y: torch.where(x <= y, x, y))(*torch.broadcast_tensors(raw(a), raw(b).to(raw(a).dtype)))))
    tf.clip_by_value = op(lambda x, lo, hi: AT(torch.clamp(raw(x), float(lo), float(hi))))
    tf.cast = op(_cast)
    tf.argmax = op(lambda x, axis=None: AT(torch.argmax(raw(x), dim=axis)))
    tf.add_n = op(lambda vals: AT(functools.reduce(torch.add, [raw(v) for v in vals])))
    tf.pad, tf.one_hot = pad, one_hot
    tf.nn = types.SimpleNamespace(
        conv2d=nn_conv2d, conv3d=nn_conv3d, softmax=nn_softmax, softmax_cross_entropy_with_logits=nn_softmax_xent, relu=relu,
        sigmoid=op(lambda x: AT(torch.sigmoid(raw(x)))), bias_add=op(lambda x, b: AT(raw(x) + raw(b))),
        l2_loss=op(lambda x: AT((raw(x) * raw(x)).sum() / 2)))

    contrib = types.ModuleType('tensorflow.contrib')
    slim = types.ModuleType('tensorflow.contrib.slim')
    slim.conv2d, slim.conv2d_transpose, slim.batch_norm = slim_conv2d, slim_conv2d_transpose, slim_batch_norm
    slim.arg_scope, slim.add_arg_scope = arg_scope, add_arg_scope
    slim.l2_regularizer = lambda s: (lambda w: AT(float(s) * (raw(w) * raw(w)).sum() / 2))
    layers = types.ModuleType('tensorflow.contrib.layers')
    layers.xavier_initializer = lambda **k: None
    contrib.slim, contrib.layers = slim, layers
    tf.contrib = contrib

    fj = types.ModuleType('fjcommon')
    th = types.ModuleType('fjcommon.tf_helpers')
    th.transpose_NHWC_to_NCHW = lambda x: tf.transpose(x, (0, 3, 1, 2))
    th.transpose_NCHW_to_NHWC = lambda x: tf.transpose(x, (0, 2, 3, 1))
    th.log10 = lambda x: AT(torch.log10(raw(x)))
    th.list_without_None = lambda *a: [v for v in a if v is not None]

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


REFERENCE_MODULES = ('autoencoder', 'quantizer', 'probclass', 'ms_ssim', 'bits')


def install(weights, dtype=torch.float64, reference_code_dir='/root/reference/code'):
    """Fresh state + shim modules in sys.modules; the reference modules are re-imported against them."""
    global _S
    _S = _State()
    _S.fdt = dtype
    _S.weights = dict(weights)
    mods = build_modules()
    sys.modules.update(mods)
    for m in REFERENCE_MODULES + ('bit_counter', 'bpp_helpers', 'ms_ssim_np', 'train'):
        sys.modules.pop(m, None)
    if reference_code_dir not in sys.path:
        sys.path.insert(0, reference_code_dir)
    return mods['tensorflow']


def uninstall():
    """Remove the shim and the reference modules imported against it (other tests install the numpy shim)."""
    for m in list(sys.modules):
        if m == 'tensorflow' or m.startswith('tensorflow.') or m == 'fjcommon' or m.startswith('fjcommon.') or \
                m in REFERENCE_MODULES + ('bit_counter', 'bpp_helpers', 'ms_ssim_np', 'train'):
            sys.modules.pop(m, None)


def load_train_definitions(tf_module, reference_code_dir='/root/reference/code'):
    """exec the UNMODIFIED source of get_loss and Distortions (code/train.py:303-336,352-431) -- train.py itself imports
    the whole experiment plumbing (input queues, savers, loggers), which is out of scope."""
    import ast
    import importlib
    path = reference_code_dir + '/train.py'
    src = open(path).read()
    tree = ast.parse(src)
    ns = {'tf': tf_module, 'np': np, 'ms_ssim': importlib.import_module('ms_ssim'),
          'tf_helpers': sys.modules['fjcommon.tf_helpers']}
    for node in tree.body:
        if (isinstance(node, ast.FunctionDef) and node.name == 'get_loss') or (isinstance(node, ast.ClassDef) and node.name == 'Distortions'):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
    return ns['get_loss'], ns['Distortions']
