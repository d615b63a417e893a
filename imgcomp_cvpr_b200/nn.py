"""Thin host wrappers of the training primitives (include/imgcomp_b200.h, ic_nn_*): torch CUDA tensors are
containers, every operation runs in libimgcomp_b200.so.  Activations NHWC float32 with channel counts that
are multiples of 4, convolution weights [KH][KW][Cin][Cout] in the orientation of the op."""
import torch

from . import _lib

BN_EPS = 1e-5        # slim.batch_norm epsilon of the reference (code/autoencoder.py:121)


def _f32(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), 'float32 contiguous CUDA tensor expected'
    return t


_ws_cache = {}


def _workspace(nbytes):
    """Process-wide scratch buffer, grown on demand.  Growing REPLACES the tensor: whoever baked its address into a CUDA
    graph must hold a reference to the old one (trainer.Trainer.enable_cuda_graph does, via current_workspace()), so the
    block cannot return to the allocator while replays still write to it."""
    ws = _ws_cache.get('ws')
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device='cuda')
        _ws_cache['ws'] = ws
    return ws


def current_workspace():
    return _ws_cache.get('ws')


def _geo(x_shape, w_shape, stride, transposed, valid):
    N, Hi, Wi, Cin = x_shape
    KH, KW, wi, Cout = w_shape
    assert wi == Cin, 'weights are [KH][KW][Cin][Cout]: Cin %d != %d' % (wi, Cin)
    if valid:
        Ho, Wo = (Hi - KH) // stride + 1, (Wi - KW) // stride + 1
    elif transposed:
        Ho, Wo = Hi * stride, Wi * stride
    else:
        Ho, Wo = -(-Hi // stride), -(-Wi // stride)
    return (N, Hi, Wi, Cin, KH, KW, stride, Cout, int(bool(transposed)), int(bool(valid))), (N, Ho, Wo, Cout)


def conv2d_fwd(x, w, stride=1, transposed=False, valid=False):
    g, oshape = _geo(x.shape, w.shape, stride, transposed, valid)
    y = torch.empty(oshape, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().ic_nn_conv2d_fwd(_lib.ptr(_f32(x)), _lib.ptr(_f32(w)), *g, _lib.ptr(y), _lib.stream_ptr()))
    return y


def conv2d_bwd_data(dy, w, x_shape, stride=1, transposed=False, valid=False, out=None):
    g, oshape = _geo(x_shape, w.shape, stride, transposed, valid)
    assert tuple(dy.shape) == oshape, '{} != {}'.format(tuple(dy.shape), oshape)
    dx = torch.empty(tuple(x_shape), dtype=torch.float32, device=dy.device) if out is None else _f32(out)
    assert tuple(dx.shape) == tuple(x_shape)
    ws = _workspace(_lib.lib().ic_nn_conv2d_workspace_bytes(*g))
    _lib.check(_lib.lib().ic_nn_conv2d_bwd_data(_lib.ptr(_f32(dy)), _lib.ptr(_f32(w)), *g, _lib.ptr(dx), _lib.ptr(ws), ws.numel(),
                                               _lib.stream_ptr()))
    return dx


def conv2d_bwd_filter(x, dy, w_shape, stride=1, transposed=False, valid=False, out=None):
    g, oshape = _geo(x.shape, w_shape, stride, transposed, valid)
    assert tuple(dy.shape) == oshape, '{} != {}'.format(tuple(dy.shape), oshape)
    dw = torch.empty(tuple(w_shape), dtype=torch.float32, device=x.device) if out is None else out
    assert tuple(dw.shape) == tuple(w_shape)
    ws = _workspace(_lib.lib().ic_nn_conv2d_workspace_bytes(*g))
    _lib.check(_lib.lib().ic_nn_conv2d_bwd_filter(_lib.ptr(_f32(x)), _lib.ptr(_f32(dy)), *g, _lib.ptr(dw), _lib.ptr(ws), ws.numel(),
                                                 _lib.stream_ptr()))
    return dw


def conv3x3_tc(x, w, data_grad=False, keep=False):
    """3x3 stride-1 SAME 128 -> 128 convolution (data_grad: its data gradient, x = dy) on the tcgen05 kernel, EXACT mode;
    x N,H,W,128 float32, w [3][3][128][128] float32 (both on the device).  keep=True: -> (y, cache) where cache holds the
    pre-scaled fp16 hi/lo planes of x and the scales of w and x for conv3x3_tc_bwd (no second maximum search / split)."""
    N, H, W, C = x.shape
    assert C == 128 and tuple(w.shape) == (3, 3, 128, 128)
    y = torch.empty_like(x)
    ws = _workspace(_lib.lib().ic_nn_conv3x3_tc_workspace_bytes(N, H, W))
    if not keep:
        _lib.check(_lib.lib().ic_nn_conv3x3_tc(_lib.ptr(_f32(x)), _lib.ptr(_f32(w)), N, H, W, int(bool(data_grad)), _lib.ptr(y),
                                              _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return y
    planes = torch.empty(2 * x.numel(), dtype=torch.float16, device=x.device)
    scales = torch.empty(4, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().ic_nn_conv3x3_tc_ex(_lib.ptr(_f32(x)), _lib.ptr(_f32(w)), N, H, W, int(bool(data_grad)), _lib.ptr(y),
                                             _lib.ptr(planes), _lib.ptr(scales), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return y, (planes, scales)


def conv3x3_tc_bwd(x, dy, w, need_dx=True, dw_out=None, cache=None):
    """backward of conv3x3_tc(x, w): -> (dx or None, dw [3][3][128][128]); both gradients on the tcgen05 kernels.
    cache: what conv3x3_tc(..., keep=True) returned for this x and w"""
    N, H, W, C = x.shape
    assert C == 128 and tuple(w.shape) == (3, 3, 128, 128) and tuple(dy.shape) == tuple(x.shape)
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w) if dw_out is None else dw_out
    assert tuple(dw.shape) == (3, 3, 128, 128)
    ws = _workspace(_lib.lib().ic_nn_conv3x3_tc_bwd_workspace_bytes(N, H, W))
    if cache is None:
        _lib.check(_lib.lib().ic_nn_conv3x3_tc_bwd(_lib.ptr(_f32(x)), _lib.ptr(_f32(dy)), _lib.ptr(_f32(w)), N, H, W, _lib.ptr(dx),
                                                  _lib.ptr(_f32(dw)), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    else:
        planes, scales = cache
        assert planes.numel() == 2 * x.numel() and planes.dtype == torch.float16
        _lib.check(_lib.lib().ic_nn_conv3x3_tc_bwd_ex(_lib.ptr(_f32(x)), _lib.ptr(_f32(dy)), _lib.ptr(_f32(w)), N, H, W, _lib.ptr(dx),
                                                     _lib.ptr(_f32(dw)), _lib.ptr(planes), _lib.ptr(scales), _lib.ptr(ws), ws.numel(),
                                                     _lib.stream_ptr()))
    return dx, dw


class TcPlan(object):
    """One conv of the training step on the tcgen05 kernels (ic_nn_tc_plan_*): op_kind 'conv5s2' (slim.conv2d 5x5 stride 2),
    'tconv5s2' (slim.conv2d_transpose 5x5 stride 2) or 'pc' (masked (2,3,3) conv3d, depth-major volume); data_grad selects
    the gradient w.r.t. the op's input.  TcPlan.get(...) -> None when no tensor-core kernel covers the shape."""
    KINDS = {'conv5s2': 0, 'tconv5s2': 1, 'pc': 3}
    _cache = {}

    def __init__(self, handle, kind, data_grad, cin, cout):
        self.h, self.kind, self.data_grad, self.cin, self.cout = handle, kind, bool(data_grad), cin, cout

    @classmethod
    def get(cls, kind, data_grad, cin, cout):
        key = (kind, bool(data_grad), cin, cout)
        if key not in cls._cache:
            import ctypes
            h = ctypes.c_void_p()
            rc = _lib.lib().ic_nn_tc_plan_create(cls.KINDS[kind], int(bool(data_grad)), cin, cout, ctypes.byref(h))
            if rc == -4:        # IC_ERR_UNSUPPORTED
                cls._cache[key] = None
            else:
                _lib.check(rc)
                cls._cache[key] = cls(h, kind, data_grad, cin, cout)
        return cls._cache[key]

    def out_shape(self, x_shape):
        c4 = lambda n: (n + 3) // 4 * 4
        co = c4(self.cin if self.data_grad else self.cout)
        if self.kind == 'pc':
            D, N, H, W, _ = x_shape
            return (D + 1, N, H + 2, W + 2, co) if self.data_grad else (D - 1, N, H - 2, W - 2, co)
        N, H, W, _ = x_shape
        strided = (self.kind == 'conv5s2') != self.data_grad
        return (N, H // 2, W // 2, co) if strided else (N, 2 * H, 2 * W, co)

    def run(self, x, w, out=None):
        """x: float32 NHWC input of this conv (dy for a data gradient); w: the op's weight buffer -> float32 NHWC output"""
        if self.kind == 'pc':
            D, N, H, W, _ = x.shape
        else:
            (N, H, W, _), D = x.shape, 1
        oshape = self.out_shape(x.shape)
        y = torch.empty(oshape, dtype=torch.float32, device=x.device) if out is None else _f32(out)
        assert tuple(y.shape) == tuple(oshape), (tuple(y.shape), oshape)
        L = _lib.lib()
        ws = _workspace(L.ic_nn_tc_plan_workspace_bytes(self.h, D, N, H, W))
        _lib.check(L.ic_nn_tc_plan_run(self.h, _lib.ptr(_f32(x)), _lib.ptr(_f32(w)), D, N, H, W, _lib.ptr(y), _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr()))
        return y


class TcWgradPlan(object):
    """Filter gradient of one of TcPlan's ops on the tcgen05 GEMM-over-pixels kernel (ic_nn_tc_wgrad_plan_*)."""
    _cache = {}

    def __init__(self, handle, kind):
        self.h, self.kind = handle, kind

    @classmethod
    def get(cls, kind, cin, cout):
        key = (kind, cin, cout)
        if key not in cls._cache:
            import ctypes
            h = ctypes.c_void_p()
            rc = _lib.lib().ic_nn_tc_wgrad_plan_create(TcPlan.KINDS[kind], cin, cout, ctypes.byref(h))
            if rc == -4:        # IC_ERR_UNSUPPORTED
                cls._cache[key] = None
            else:
                _lib.check(rc)
                cls._cache[key] = cls(h, kind)
        return cls._cache[key]

    def run(self, x, dy, dw_out):
        """x: the op's input, dy: gradient w.r.t. its output (float32 NHWC; depth-major 5-D for 'pc') -> dw_out (filled)"""
        if self.kind == 'pc':
            D, N, H, W, _ = x.shape
        else:
            (N, H, W, _), D = x.shape, 1
        L = _lib.lib()
        ws = _workspace(L.ic_nn_tc_wgrad_plan_workspace_bytes(self.h, D, N, H, W))
        _lib.check(L.ic_nn_tc_wgrad_plan_run(self.h, _lib.ptr(_f32(x)), _lib.ptr(_f32(dy)), D, N, H, W, _lib.ptr(_f32(dw_out)), _lib.ptr(ws),
                                             ws.numel(), _lib.stream_ptr()))
        return dw_out


def bn_train_fwd(x, gamma, beta, relu=False, res1=None, res2=None, mov_mean=None, mov_var=None, stats=None, partial=None,
                 planes_out=None, write_out=True):
    """x (..., C) -> (out, mean, invstd).  stats = (mean, invstd): plain affine layer with the given statistics.
    partial: statistics partial sums of x from conv3x3_tc_fused (no statistics pass over x); planes_out (x N,H,W,C with
    C % 8 == 0): float16 tensor of 2 * x.numel() elements that receives the hi/lo planes of out for the next 3x3 conv;
    write_out=False (with planes_out): `out` is returned as an UNWRITTEN placeholder (only the planes are produced)."""
    assert write_out or planes_out is not None
    C = x.shape[-1]
    M = x.numel() // C
    out = torch.empty_like(x)
    if stats is None:
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
    else:
        mean, invstd = stats
    ws = _workspace(_lib.lib().ic_nn_bn_workspace_bytes(M, C))
    hw = 0
    if planes_out is not None:
        assert x.dim() == 4 and planes_out.dtype == torch.float16 and planes_out.numel() == 2 * x.numel()
        hw = x.shape[1] * x.shape[2]
    _lib.check(_lib.lib().ic_nn_bn_train_fwd_ex(_lib.ptr(_f32(x)), M, C, _lib.ptr(_f32(gamma)), _lib.ptr(_f32(beta)), BN_EPS,
                                               int(relu), int(stats is None), _lib.ptr(res1), _lib.ptr(res2), _lib.ptr(mean),
                                               _lib.ptr(invstd), _lib.ptr(mov_mean), _lib.ptr(mov_var), _lib.ptr(out if write_out else None),
                                               _lib.ptr(partial if stats is None else None), _lib.ptr(planes_out), hw,
                                               _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return out, mean, invstd


def weight_scales(base, offsets, count, out):
    """power-of-two scales of the len(offsets) weight tensors base[offsets[i] : offsets[i] + count] -> out (n, 4) float32:
    row i = (scale, 1, 1 / scale, 1), the `scales` of conv3x3_tc_bwd's cache.  One launch."""
    assert offsets.dtype == torch.int64 and out.shape == (offsets.numel(), 4)
    _lib.check(_lib.lib().ic_nn_weight_scales(_lib.ptr(_f32(base)), _lib.ptr(offsets), offsets.numel(), count, _lib.ptr(_f32(out)),
                                             _lib.stream_ptr()))
    return out


def bn_partial_buffer(M, device):
    return torch.empty(_lib.lib().ic_nn_bn_partial_bytes(M) // 8, dtype=torch.float64, device=device)


def pack3x3_all(base, offsets, scales, W, out=None):
    """packed weight images of all trunk convs (forward + data gradient) in one launch -> uint8 tensor (n, 2, entry bytes);
    row [i, 0] / [i, 1] is the `prepared` argument of conv3x3_tc_fused / conv3x3_tc_bwd_planes for tensor i"""
    L = _lib.lib()
    n, eb = offsets.numel(), L.ic_nn_conv3x3_tc_prepared_bytes()
    if out is None:
        out = torch.empty((n, 2, eb), dtype=torch.uint8, device=base.device)
    assert out.shape == (n, 2, eb) and out.is_contiguous()
    _lib.check(L.ic_nn_pack3x3_all(_lib.ptr(_f32(base)), _lib.ptr(offsets), _lib.ptr(_f32(scales)), n, int(W), _lib.ptr(out), _lib.stream_ptr()))
    return out


def conv3x3_tc_fused(x_planes, w, wscale, shape, partial, prepared=None):
    """conv3x3_tc for an input that already exists as UNSCALED fp16 hi/lo planes (bn_train_fwd(planes_out=...)): no maximum
    search, no split; -> y (float32 NHWC) and the batch-norm partial sums of y in `partial` (bn_partial_buffer)."""
    N, H, W, C = shape
    assert C == 128 and tuple(w.shape) == (3, 3, 128, 128) and x_planes.numel() == 2 * N * H * W * C
    y = torch.empty(shape, dtype=torch.float32, device=w.device)
    L = _lib.lib()
    assert partial.numel() * 8 >= L.ic_nn_bn_partial_bytes(N * H * W)
    ws = _workspace(L.ic_nn_conv3x3_tc_fused_workspace_bytes(N, H, W))
    _lib.check(L.ic_nn_conv3x3_tc_fused(_lib.ptr(x_planes), _lib.ptr(_f32(w)), _lib.ptr(_f32(wscale)), _lib.ptr(prepared), N, H, W, _lib.ptr(y), _lib.ptr(partial),
                                        _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return y


def bn_train_bwd(x, dy, gamma, beta, mean, invstd, relu=False, use_stats=True, dgamma=None, dbeta=None):
    """-> (dx, dgamma, dbeta)"""
    C = x.shape[-1]
    M = x.numel() // C
    dx = torch.empty_like(x)
    dgamma = torch.empty(C, dtype=torch.float32, device=x.device) if dgamma is None else dgamma
    dbeta = torch.empty(C, dtype=torch.float32, device=x.device) if dbeta is None else dbeta
    ws = _workspace(_lib.lib().ic_nn_bn_workspace_bytes(M, C))
    _lib.check(_lib.lib().ic_nn_bn_train_bwd(_lib.ptr(_f32(x)), _lib.ptr(_f32(dy)), M, C, _lib.ptr(_f32(gamma)), _lib.ptr(_f32(beta)),
                                            int(relu), int(use_stats), _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(dx),
                                            _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return dx, dgamma, dbeta


class PlanesGrad(object):
    """An output gradient that exists only as pre-scaled fp16 hi/lo planes (bn_train_bwd(..., as_planes=True)): what the
    tensor-core backward of the conv before the batch norm consumes (conv3x3_tc_bwd_planes)."""

    def __init__(self, planes, scale, shape):
        self.planes, self.scale, self.shape = planes, scale, tuple(shape)


def bn_train_bwd_planes(x, dy, gamma, beta, mean, invstd, relu=False, dgamma=None, dbeta=None):
    """bn_train_bwd (batch statistics, x N,H,W,128) whose dx is written as fp16 planes scaled by a power of two ->
    (PlanesGrad, dgamma, dbeta); no float32 dx exists."""
    N, H, W, C = x.shape
    assert C == 128
    M = N * H * W
    planes = torch.empty(2 * x.numel(), dtype=torch.float16, device=x.device)
    scale = torch.empty(2, dtype=torch.float32, device=x.device)
    dgamma = torch.empty(C, dtype=torch.float32, device=x.device) if dgamma is None else dgamma
    dbeta = torch.empty(C, dtype=torch.float32, device=x.device) if dbeta is None else dbeta
    ws = _workspace(_lib.lib().ic_nn_bn_workspace_bytes(M, C))
    _lib.check(_lib.lib().ic_nn_bn_train_bwd_ex(_lib.ptr(_f32(x)), _lib.ptr(_f32(dy)), M, C, _lib.ptr(_f32(gamma)), _lib.ptr(_f32(beta)),
                                               int(relu), 1, _lib.ptr(mean), _lib.ptr(invstd), None, _lib.ptr(dgamma), _lib.ptr(dbeta),
                                               _lib.ptr(planes), _lib.ptr(scale), H * W, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return PlanesGrad(planes, scale, x.shape), dgamma, dbeta


def conv3x3_tc_bwd_planes(dy, w, need_dx=True, dw_out=None, cache=None, dx_add=None, prepared=None):
    """conv3x3_tc_bwd for dy given as a PlanesGrad; cache = (x planes, scales) of the forward pass (required);
    dx_add: gradient the input already has (residual path) -> dx = data gradient + dx_add in the same pass"""
    N, H, W, C = dy.shape
    planes, scales = cache
    assert C == 128 and planes.numel() == dy.planes.numel()
    dx = torch.empty(dy.shape, dtype=torch.float32, device=w.device) if need_dx else None
    dw = torch.empty_like(w) if dw_out is None else dw_out
    ws = _workspace(_lib.lib().ic_nn_conv3x3_tc_bwd_workspace_bytes(N, H, W))
    assert dx_add is None or (need_dx and tuple(dx_add.shape) == tuple(dy.shape))
    _lib.check(_lib.lib().ic_nn_conv3x3_tc_bwd_planes(_lib.ptr(dy.planes), _lib.ptr(dy.scale), _lib.ptr(_f32(w)), _lib.ptr(prepared), N, H, W, _lib.ptr(dx),
                                                     _lib.ptr(None if dx_add is None else _f32(dx_add)), _lib.ptr(_f32(dw)), _lib.ptr(planes),
                                                     _lib.ptr(scales), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return dx, dw


def normalize_fwd(x):
    """x N,3,H,W uint8 or float32 in [0,255] -> normalised N,H,W,4 (code/autoencoder.py:136-144)"""
    assert x.is_cuda and x.is_contiguous() and x.dim() == 4 and x.shape[1] == 3 and x.dtype in (torch.uint8, torch.float32)
    N, _, H, W = x.shape
    out = torch.empty((N, H, W, 4), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().ic_nn_normalize_fwd(_lib.ptr(x), int(x.dtype == torch.uint8), N, H, W, _lib.ptr(out), _lib.stream_ptr()))
    return out


def hq_fwd(bn, C, heatmap, centers):
    """bn N,h,w,Cb -> dict of NCHW tensors z, heatmap, qbar, qhard, qsoft (float32) and symbols (int64)"""
    N, h, w, Cb = bn.shape
    o = {k: torch.empty((N, C, h, w), dtype=torch.float32, device=bn.device) for k in ('z', 'heatmap', 'qbar', 'qhard', 'qsoft')}
    o['symbols'] = torch.empty((N, C, h, w), dtype=torch.int64, device=bn.device)
    _lib.check(_lib.lib().ic_nn_hq_fwd(_lib.ptr(_f32(bn)), N, h, w, C, Cb, int(heatmap), _lib.ptr(_f32(centers)), centers.numel(),
                                      _lib.ptr(o['z']), _lib.ptr(o['heatmap']), _lib.ptr(o['qbar']), _lib.ptr(o['qhard']),
                                      _lib.ptr(o['qsoft']), _lib.ptr(o['symbols']), _lib.stream_ptr()))
    if not heatmap:
        o['heatmap'] = None
    return o


def hq_bwd(bn, C, heatmap, centers, dq, dhm=None, dcenters=None):
    """bn N,h,w,Cb; dq N,h,w,C; dhm N,C,h,w or None -> (dbn N,h,w,Cb, dcenters (L,))"""
    N, h, w, Cb = bn.shape
    dbn = torch.empty_like(bn)
    dcent = torch.empty(centers.numel(), dtype=torch.float32, device=bn.device) if dcenters is None else dcenters
    ws = _workspace(_lib.lib().ic_nn_hq_workspace_bytes(N * h * w))
    _lib.check(_lib.lib().ic_nn_hq_bwd(_lib.ptr(_f32(bn)), N, h, w, C, Cb, int(heatmap), _lib.ptr(_f32(centers)), centers.numel(),
                                      _lib.ptr(_f32(dq)), _lib.ptr(dhm), _lib.ptr(dbn), _lib.ptr(dcent), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr()))
    return dbn, dcent


def denorm_clip_fwd(v):
    """v N,H,W,4 -> x_out N,3,H,W in [0,255]"""
    N, H, W, _ = v.shape
    out = torch.empty((N, 3, H, W), dtype=torch.float32, device=v.device)
    _lib.check(_lib.lib().ic_nn_denorm_clip_fwd(_lib.ptr(_f32(v)), N, H, W, _lib.ptr(out), _lib.stream_ptr()))
    return out


def denorm_clip_bwd(v, dx_out):
    N, H, W, _ = v.shape
    dv = torch.empty_like(v)
    _lib.check(_lib.lib().ic_nn_denorm_clip_bwd(_lib.ptr(_f32(v)), _lib.ptr(_f32(dx_out)), N, H, W, _lib.ptr(dv), _lib.stream_ptr()))
    return dv


def nhwc_to_nchw(x, C=None):
    N, H, W, Cs = x.shape
    C = Cs if C is None else C
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().ic_nn_nhwc_to_nchw(_lib.ptr(_f32(x)), N, C, Cs, H * W, _lib.ptr(out), _lib.stream_ptr()))
    return out


def nchw_to_nhwc(x, Cs=None):
    N, C, H, W = x.shape
    Cs = C if Cs is None else Cs
    out = torch.empty((N, H, W, Cs), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().ic_nn_nchw_to_nhwc(_lib.ptr(_f32(x)), N, C, Cs, H * W, _lib.ptr(out), _lib.stream_ptr()))
    return out


def axpby(a, x, b=0.0, y=None, out=None):
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.lib().ic_nn_axpby(float(a), _lib.ptr(_f32(x)), float(b), _lib.ptr(y), x.numel(), _lib.ptr(out), _lib.stream_ptr()))
    return out


def add(x, y):
    return axpby(1.0, x, 1.0, y)


def mul(x, y, out=None):
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.lib().ic_nn_mul(_lib.ptr(_f32(x)), _lib.ptr(_f32(y)), x.numel(), _lib.ptr(out), _lib.stream_ptr()))
    return out


def adam_step(w, grad, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, l2=0.0, mask=None):
    """in place on w, m, v"""
    _lib.check(_lib.lib().ic_nn_adam_step(_lib.ptr(_f32(w)), _lib.ptr(_f32(grad)), _lib.ptr(_f32(m)), _lib.ptr(_f32(v)), w.numel(),
                                         float(lr), float(beta1), float(beta2), float(eps), int(step), float(l2), _lib.ptr(mask),
                                         _lib.stream_ptr()))


def pc_pad_fwd(q, pad_value):
    """pad_for_probclass3d (code/probclass.py:268-292): q N,C,h,w -> depth-major (C+4, N, h+8, w+8, 4), channel 0 = value.
    pad_value: host float, or a CUDA tensor whose first element is read on the device (no host read-back)"""
    N, C, h, w = q.shape
    out = torch.empty((C + 4, N, h + 8, w + 8, 4), dtype=torch.float32, device=q.device)
    dev = pad_value if torch.is_tensor(pad_value) else None
    _lib.check(_lib.lib().ic_nn_pc_pad_fwd(_lib.ptr(_f32(q)), N, C, h, w, 0.0 if dev is not None else float(pad_value), _lib.ptr(dev),
                                          _lib.ptr(out), _lib.stream_ptr()))
    return out


def pc_xent_fwd(logits, L, symbols):
    """logits (C, N, h, w, Cs) -> bits N,C,h,w (code/probclass.py:99-104)"""
    C, N, h, w, Cs = logits.shape
    assert symbols.dtype == torch.int64 and symbols.is_contiguous() and tuple(symbols.shape) == (N, C, h, w)
    bc = torch.empty((N, C, h, w), dtype=torch.float32, device=logits.device)
    _lib.check(_lib.lib().ic_nn_pc_xent_fwd(_lib.ptr(_f32(logits)), Cs, L, _lib.ptr(symbols), N, C, h, w, _lib.ptr(bc), _lib.stream_ptr()))
    return bc


def pc_xent_bwd(logits, L, symbols, heatmap, coef_real=0.0, coef_mask=0.0, coef_dev=None):
    """-> d logits (C, N, h, w, Cs) for the upstream gradient (coef_real + coef_mask * heatmap) per symbol;
    coef_dev (1-element CUDA tensor from rate_coef) overrides both coefficients on the device"""
    C, N, h, w, Cs = logits.shape
    d = torch.empty_like(logits)
    _lib.check(_lib.lib().ic_nn_pc_xent_bwd(_lib.ptr(_f32(logits)), Cs, L, _lib.ptr(symbols), _lib.ptr(heatmap), N, C, h, w,
                                           float(coef_real), float(coef_mask), _lib.ptr(coef_dev), _lib.ptr(d), _lib.stream_ptr()))
    return d


def rate_coef(sums, n, beta, h_target, has_heatmap, out):
    """out[0] = beta * 0.5 / n if the rate hinge of code/train.py:309-316 is active else 0, computed on the device"""
    _lib.check(_lib.lib().ic_nn_rate_coef(_lib.ptr(sums), int(n), float(beta), float(h_target), int(bool(has_heatmap)), _lib.ptr(out),
                                         _lib.stream_ptr()))
    return out


def scale_dev(a, x):
    """a[0] * x with the scalar in device memory"""
    out = torch.empty_like(x)
    _lib.check(_lib.lib().ic_nn_scale_dev(_lib.ptr(a), _lib.ptr(_f32(x)), x.numel(), _lib.ptr(out), _lib.stream_ptr()))
    return out


def adam_step_dev(w, grad, m, v, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, l2=0.0, mask=None):
    """in place on w, m, v; lr_t: 1-element CUDA tensor holding lr * sqrt(1 - beta2^t) / (1 - beta1^t)"""
    _lib.check(_lib.lib().ic_nn_adam_step_dev(_lib.ptr(_f32(w)), _lib.ptr(_f32(grad)), _lib.ptr(_f32(m)), _lib.ptr(_f32(v)), w.numel(),
                                             _lib.ptr(lr_t), float(beta1), float(beta2), float(eps), float(l2), _lib.ptr(mask),
                                             _lib.stream_ptr()))


def crop_fwd(x, crop):
    """x (..., H, W, C) -> x[..., crop:-crop, crop:-crop, :]"""
    H, W, C = x.shape[-3:]
    A = x.numel() // (H * W * C)
    out = torch.empty(tuple(x.shape[:-3]) + (H - 2 * crop, W - 2 * crop, C), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().ic_nn_crop_fwd(_lib.ptr(_f32(x)), A, H, W, C, crop, _lib.ptr(out), _lib.stream_ptr()))
    return out


def crop_bwd_add(dy, dx, crop):
    """dx[..., crop:-crop, crop:-crop, :] += dy (in place)"""
    H, W, C = dx.shape[-3:]
    A = dx.numel() // (H * W * C)
    assert dy.numel() == A * (H - 2 * crop) * (W - 2 * crop) * C
    _lib.check(_lib.lib().ic_nn_crop_bwd_add(_lib.ptr(_f32(dy)), A, H, W, C, crop, _lib.ptr(_f32(dx)), _lib.stream_ptr()))
    return dx


def msssim_tf_bwd(img1, img2, grad_out):
    """gradient of ms_ssim.MultiScaleSSIM(img1, img2) w.r.t. img2, times grad_out -> (d img2, value (1,) tensor)"""
    assert img1.shape == img2.shape and img1.dim() == 4 and img1.shape[1] == 3
    N, _, H, W = img1.shape
    d = torch.empty_like(img2)
    val = torch.empty(1, dtype=torch.float32, device=img1.device)
    ws = _workspace(_lib.lib().ic_msssim_bwd_workspace_bytes(N, H, W))
    _lib.check(_lib.lib().ic_msssim_tf_bwd(_lib.ptr(_f32(img1)), _lib.ptr(_f32(img2)), N, H, W, float(grad_out), _lib.ptr(d),
                                          _lib.ptr(val), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return d, val


def distortion_bwd(x, x_out, psnr=False):
    """distortion_to_minimize 'mse' / 'psnr' (code/train.py:381-397, float32 branch): -> (d loss / d x_out, per-image MSE)"""
    N = x.shape[0]
    per = x[0].numel()
    mse = torch.empty(N, dtype=torch.float32, device=x.device)
    d = torch.empty_like(x_out)
    L = _lib.lib()
    _lib.check(L.ic_mse_per_image_fwd(_lib.ptr(_f32(x)), _lib.ptr(_f32(x_out)), N, per, 0, _lib.ptr(mse), _lib.stream_ptr()))
    _lib.check(L.ic_nn_distortion_bwd(_lib.ptr(x), _lib.ptr(x_out), _lib.ptr(mse), N, per, int(bool(psnr)), _lib.ptr(d), _lib.stream_ptr()))
    return d, mse
