"""Training primitives (ic_nn_*) against torch-CPU float64 autograd of the oracle's own building blocks
(oracle/train_oracle.py): each backward kernel must be the gradient of what the forward kernel computes,
with TF's SAME padding / conv2d_transpose semantics."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import train_oracle as T

pytestmark = pytest.mark.gpu


def _rnd(shape, seed, scale=1.0):
    return (np.random.RandomState(seed).randn(*shape) * scale).astype(np.float32)


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _nhwc(t):       # NCHW torch -> NHWC numpy
    return t.detach().permute(0, 2, 3, 1).contiguous().numpy()


CONV_CASES = [
    # N, H, W, Cin, Cout, K, stride, transposed
    (2, 9, 11, 4, 8, 3, 1, False),
    (2, 10, 12, 4, 64, 5, 2, False),       # h1-like: even size -> pad (1, 2)
    (1, 9, 7, 8, 12, 5, 2, False),         # odd size
    (2, 6, 5, 8, 4, 3, 2, True),           # from_bn-like
    (1, 5, 6, 12, 8, 5, 2, True),          # h12-like
    (3, 8, 8, 128, 128, 3, 1, False),      # residual conv
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_fwd_bwd(case):
    from imgcomp_cvpr_b200 import nn
    N, H, W, Cin, Cout, K, s, tr = case
    x = _rnd((N, Cin, H, W), 1)
    w = _rnd((K, K, Cin, Cout), 2, 0.2)                       # op orientation
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    if tr:
        y = T.conv2d_transpose_same(xt, wt.permute(0, 1, 3, 2), s)      # TF variable is [kh, kw, out, in]
    else:
        y = T.conv2d_same(xt, wt, s)
    dy = _rnd(tuple(y.shape), 3)
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    xg = _cuda(x.transpose(0, 2, 3, 1))
    wg = _cuda(w)
    yg = nn.conv2d_fwd(xg, wg, s, tr)
    np.testing.assert_allclose(yg.cpu().numpy(), _nhwc(y), rtol=2e-5, atol=2e-5 * np.abs(_nhwc(y)).max())
    dyg = _cuda(dy.transpose(0, 2, 3, 1))
    dx = nn.conv2d_bwd_data(dyg, wg, xg.shape, s, tr)
    ref = _nhwc(xt.grad)
    np.testing.assert_allclose(dx.cpu().numpy(), ref, rtol=2e-5, atol=2e-5 * np.abs(ref).max())
    dw = nn.conv2d_bwd_filter(xg, dyg, wg.shape, s, tr)
    ref = wt.grad.numpy()
    np.testing.assert_allclose(dw.cpu().numpy(), ref, rtol=3e-5, atol=3e-5 * np.abs(ref).max())


def test_conv_valid_mode():
    from imgcomp_cvpr_b200 import nn
    x, w = _rnd((2, 4, 9, 8), 4), _rnd((3, 3, 4, 8), 5, 0.3)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xt, wt.permute(3, 2, 0, 1))
    dy = _rnd(tuple(y.shape), 6)
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    xg, wg, dyg = _cuda(x.transpose(0, 2, 3, 1)), _cuda(w), _cuda(dy.transpose(0, 2, 3, 1))
    np.testing.assert_allclose(nn.conv2d_fwd(xg, wg, valid=True).cpu().numpy(), _nhwc(y), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(nn.conv2d_bwd_data(dyg, wg, xg.shape, valid=True).cpu().numpy(), _nhwc(xt.grad), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(nn.conv2d_bwd_filter(xg, dyg, wg.shape, valid=True).cpu().numpy(), wt.grad.numpy(), rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('C,relu,res', [(64, True, False), (128, False, True), (36, False, False), (4, True, True)])
def test_batch_norm_train(C, relu, res):
    from imgcomp_cvpr_b200 import nn
    N, H, W = 3, 7, 5
    x = _rnd((N, C, H, W), 7, 2.0) + 0.5
    g, b = _rnd((C,), 8, 0.3) + 1.0, _rnd((C,), 9, 0.3)
    r = _rnd((N, C, H, W), 10)
    mm, mv = _rnd((C,), 11), np.abs(_rnd((C,), 12)) + 0.5
    P = {'s/BatchNorm/gamma': torch.tensor(g, dtype=torch.float64, requires_grad=True),
         's/BatchNorm/beta': torch.tensor(b, dtype=torch.float64, requires_grad=True)}
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    stats = {}
    y = T.batch_norm(xt, P, 's', True, stats)
    if relu:
        y = torch.relu(y)
    if res:
        y = y + torch.tensor(r, dtype=torch.float64)
    dy = _rnd((N, C, H, W), 13)
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    xg, gg, bg = _cuda(x.transpose(0, 2, 3, 1)), _cuda(g), _cuda(b)
    mmg, mvg = _cuda(mm), _cuda(mv)
    out, mean, invstd = nn.bn_train_fwd(xg, gg, bg, relu, _cuda(r.transpose(0, 2, 3, 1)) if res else None, None, mmg, mvg)
    np.testing.assert_allclose(out.cpu().numpy(), _nhwc(y), rtol=2e-5, atol=2e-5)
    mu, unb = stats['s']
    np.testing.assert_allclose(mmg.cpu().numpy(), 0.9 * mm + 0.1 * mu.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mvg.cpu().numpy(), 0.9 * mv + 0.1 * unb.numpy(), rtol=1e-5, atol=1e-6)
    dx, dgamma, dbeta = nn.bn_train_bwd(xg, _cuda(dy.transpose(0, 2, 3, 1)), gg, bg, mean, invstd, relu)
    np.testing.assert_allclose(dx.cpu().numpy(), _nhwc(xt.grad), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(dgamma.cpu().numpy(), P['s/BatchNorm/gamma'].grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dbeta.cpu().numpy(), P['s/BatchNorm/beta'].grad.numpy(), rtol=1e-4, atol=1e-4)


def test_batch_norm_statistics_with_large_mean():
    """|mean| >> std in a channel (never the case with the synthetic weights): E[x^2] - mean^2 on float32 sums would
    lose the variance to cancellation; the kernel sums (x - pivot) and (x - pivot)^2 around a data point instead"""
    from imgcomp_cvpr_b200 import nn
    rng = np.random.RandomState(0)
    M, C = 4096, 64
    x = (rng.standard_normal((M, C)) * 0.05 + 100.0 * (1 + np.arange(C))[None, :]).astype(np.float32)
    xg = _cuda(x.reshape(1, 64, 64, C))
    ones, zeros = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
    out, mean, invstd = nn.bn_train_fwd(xg, ones, zeros, False)
    x64 = x.astype(np.float64)
    np.testing.assert_allclose(mean.cpu().numpy(), x64.mean(0), rtol=1e-7)
    np.testing.assert_allclose(invstd.cpu().numpy(), 1.0 / np.sqrt(x64.var(0) + 1e-5), rtol=2e-3)
    np.testing.assert_allclose(out.cpu().numpy().reshape(M, C), (x64 - x64.mean(0)) / np.sqrt(x64.var(0) + 1e-5), atol=2e-2)


def test_affine_relu_layer():
    """use_stats = 0: y = relu(x + bias), the bias + ReLU of the context model's conv3d (code/probclass.py:259-261)"""
    from imgcomp_cvpr_b200 import nn
    M, C = 300, 24
    x, bias, dy = _rnd((M, C), 14), _rnd((C,), 15), _rnd((M, C), 16)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
    (torch.relu(xt + bt) * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    ones, zeros = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
    out, _, _ = nn.bn_train_fwd(_cuda(x), ones, _cuda(bias), True, stats=(zeros, ones))
    np.testing.assert_allclose(out.cpu().numpy(), np.maximum(x + bias, 0), rtol=1e-6, atol=1e-6)
    dx, _, dbeta = nn.bn_train_bwd(_cuda(x), _cuda(dy), ones, _cuda(bias), zeros, ones, True, use_stats=False)
    np.testing.assert_allclose(dx.cpu().numpy(), xt.grad.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(dbeta.cpu().numpy(), bt.grad.numpy(), rtol=1e-5, atol=1e-5)


def test_heatmap_quantizer_backward():
    from imgcomp_cvpr_b200 import nn
    N, C, h, w, Cb = 2, 32, 5, 6, 36
    bn = _rnd((N, C + 1, h, w), 17, 1.5)
    bn[:, 0] = _rnd((N, h, w), 18, 1.0)                          # heatmap logits around 0 -> partial masks
    centers = np.array([-1.6, -0.9, -0.2, 0.4, 1.1, 1.8], np.float32)
    dq, dhm = _rnd((N, C, h, w), 19), _rnd((N, C, h, w), 20, 0.1)
    bt = torch.tensor(bn, dtype=torch.float64, requires_grad=True)
    ct = torch.tensor(centers, dtype=torch.float64, requires_grad=True)
    hm2d = torch.sigmoid(bt[:, 0]) * C
    hm = torch.clamp(hm2d[:, None] - torch.arange(C, dtype=torch.float64).reshape(1, C, 1, 1), 0, 1)
    z = hm * bt[:, 1:]
    dist = (z[..., None] - ct) ** 2
    qsoft = (torch.softmax(-dist, -1) * ct).sum(-1)
    ((qsoft * torch.tensor(dq, dtype=torch.float64)).sum() + (hm * torch.tensor(dhm, dtype=torch.float64)).sum()).backward()
    bn_nhwc = np.zeros((N, h, w, Cb), np.float32)
    bn_nhwc[..., :C + 1] = bn.transpose(0, 2, 3, 1)
    dbn, dcent = nn.hq_bwd(_cuda(bn_nhwc), C, True, _cuda(centers), _cuda(dq.transpose(0, 2, 3, 1)), _cuda(dhm))
    ref = _nhwc(bt.grad)
    np.testing.assert_allclose(dbn.cpu().numpy()[..., :C + 1], ref, rtol=2e-4, atol=2e-5 * np.abs(ref).max())
    assert np.all(dbn.cpu().numpy()[..., C + 1:] == 0)
    np.testing.assert_allclose(dcent.cpu().numpy(), ct.grad.numpy(), rtol=2e-4, atol=1e-4)


def test_denorm_clip_and_layouts():
    from imgcomp_cvpr_b200 import nn
    from oracle import imgcomp_oracle as O
    v = _rnd((2, 6, 5, 4), 21, 2.0)
    dout = _rnd((2, 3, 6, 5), 22)
    vt = torch.tensor(v[..., :3].transpose(0, 3, 1, 2), dtype=torch.float64, requires_grad=True)
    mul = torch.tensor(np.sqrt(O.NORM_VAR + np.float32(1e-10)).astype(np.float64))[None, :, None, None]
    mean = torch.tensor(O.NORM_MEAN.astype(np.float64))[None, :, None, None]
    out = torch.clamp(vt * mul + mean, 0, 255)
    (out * torch.tensor(dout, dtype=torch.float64)).sum().backward()
    np.testing.assert_allclose(nn.denorm_clip_fwd(_cuda(v)).cpu().numpy(), out.detach().numpy(), rtol=1e-6, atol=1e-4)
    dv = nn.denorm_clip_bwd(_cuda(v), _cuda(dout)).cpu().numpy()
    np.testing.assert_allclose(dv[..., :3], _nhwc(vt.grad), rtol=1e-6, atol=1e-5)
    assert np.all(dv[..., 3] == 0)
    x = _rnd((2, 33, 4, 3), 23)
    nhwc = nn.nchw_to_nhwc(_cuda(x), 36)
    assert np.array_equal(nhwc.cpu().numpy()[..., :33], x.transpose(0, 2, 3, 1)) and np.all(nhwc.cpu().numpy()[..., 33:] == 0)
    assert np.array_equal(nn.nhwc_to_nchw(nhwc, 33).cpu().numpy(), x)


def test_adam_and_axpby():
    from imgcomp_cvpr_b200 import nn
    w, g = _rnd((1000,), 24), _rnd((1000,), 25)
    m, v = np.zeros(1000, np.float32), np.zeros(1000, np.float32)
    wg, mg, vg = _cuda(w), _cuda(m), _cuda(v)
    wr, mr, vr = w.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for step in (1, 2, 3):
        nn.adam_step(wg, _cuda(g), mg, vg, 8e-5, step, l2=0.005)
        wr, mr, vr = T.adam_update(wr, g.astype(np.float64) + 0.005 * wr, mr, vr, step, 8e-5)
    np.testing.assert_allclose(wg.cpu().numpy(), wr, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(nn.axpby(2.0, _cuda(w), -1.0, _cuda(g)).cpu().numpy(), 2 * w - g, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(nn.mul(_cuda(w), _cuda(g)).cpu().numpy(), w * g, rtol=1e-6)


# ---- planned tensor-core convs (ic_nn_tc_plan_*): forward and data gradient against float64 autograd
def _rel(a, ref):
    return float(np.abs(a - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize('case', [
    # op kind, N, H, W, Cin, Cout, gradient scale
    ('conv5s2', 2, 24, 40, 64, 128, 1.0),        # h2
    ('conv5s2', 1, 16, 16, 64, 128, 3e-5),       # tiny gradients: the per-tensor power-of-two pre-scale keeps float32-class precision
    ('tconv5s2', 2, 12, 20, 128, 64, 1.0),       # h12
    ('tconv5s2', 3, 10, 8, 128, 64, 2e-4),
    ('tconv5s2', 2, 20, 36, 64, 3, 1.0),         # h13 (forward only: its data gradient has 3 input channels)
    ('conv5s2', 2, 24, 40, 128, 33, 1.0),        # to_bn, cvpr/low (forward only: 48 columns, float32 output from the kernel)
    ('conv5s2', 1, 16, 32, 128, 65, 1.0),        # to_bn, cvpr/hi (80 columns)
])
def test_tc_plan_strided_convs(case):
    from imgcomp_cvpr_b200 import nn
    kind, N, H, W, Cin, Cout, gs = case
    tr = kind == 'tconv5s2'
    x = _rnd((N, Cin, H, W), 11)
    w = _rnd((5, 5, Cin, Cout), 12, 0.05)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    y = T.conv2d_transpose_same(xt, wt.permute(0, 1, 3, 2), 2) if tr else T.conv2d_same(xt, wt, 2)
    dy = _rnd(tuple(y.shape), 13, gs)
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    c4 = lambda n: (n + 3) // 4 * 4
    wp = np.zeros((5, 5, c4(Cin), c4(Cout)), np.float32)
    wp[:, :, :Cin, :Cout] = w
    xg, wg = _cuda(x.transpose(0, 2, 3, 1)), _cuda(wp)
    pf, pd = nn.TcPlan.get(kind, False, Cin, Cout), nn.TcPlan.get(kind, True, Cin, Cout)
    assert pf is not None
    yg = pf.run(xg, wg).cpu().numpy()
    ref = _nhwc(y)
    assert yg.shape[-1] == c4(Cout)
    err = _rel(yg[..., :Cout], ref)
    print('tc plan %s fwd %d->%d: rel err %.2e' % (kind, Cin, Cout, err))
    assert err < 2e-5
    assert not yg[..., Cout:].any()
    pw = nn.TcWgradPlan.get(kind, Cin, Cout)
    if Cout in (3, 33, 65):
        assert pd is None and pw is None
        return
    assert pd is not None
    dyp = np.zeros(tuple(ref.shape[:3]) + (c4(Cout),), np.float32)
    dyp[..., :Cout] = dy.transpose(0, 2, 3, 1)
    dx = pd.run(_cuda(dyp), wg).cpu().numpy()
    refd = _nhwc(xt.grad)
    err = _rel(dx[..., :Cin], refd)
    print('tc plan %s dgrad: rel err %.2e (gradient scale %g)' % (kind, err, gs))
    assert err < 2e-5
    # and against the FFMA kernels the trainer used before
    y32 = nn.conv2d_fwd(xg, wg, 2, tr).cpu().numpy()
    assert _rel(yg, y32) < 2e-5
    # filter gradient: GEMM over pixels on the space-to-depth form of the finer tensor
    assert pw is not None
    dw = pw.run(xg, _cuda(dyp), torch.full((5, 5, c4(Cin), c4(Cout)), 7.0, device='cuda')).cpu().numpy()
    err = _rel(dw[:, :, :Cin, :Cout], wt.grad.numpy())
    print('tc plan %s wgrad: rel err %.2e' % (kind, err))
    assert err < 2e-5
    dw32 = nn.conv2d_bwd_filter(xg, _cuda(dyp), wg.shape, 2, tr).cpu().numpy()
    assert _rel(dw, dw32) < 2e-5


@pytest.mark.parametrize('case', [
    # D, N, H, W, Cin, Cout, gradient scale
    (7, 3, 13, 11, 24, 24, 1.0),
    (5, 2, 30, 21, 24, 24, 1e-5),
    (6, 2, 12, 19, 24, 6, 1.0),
    (2, 1, 3, 3, 24, 6, 1.0),                    # one output voxel per image
])
def test_tc_plan_context_model_layers(case):
    """masked (2,3,3) VALID conv3d on the depth-major volume, forward + data gradient, against float64 autograd of
    F.conv3d with the "other" mask (code/probclass.py:164-176)"""
    from imgcomp_cvpr_b200 import nn
    D, N, H, W, Ci, Co, gs = case
    c4 = lambda n: (n + 3) // 4 * 4
    mask = np.ones((2, 3, 3), np.float32)
    mask[1, 1, 2] = 0
    mask[1, 2, :] = 0
    x = _rnd((D, N, H, W, Ci), 21)
    w = _rnd((2, 3, 3, Ci, Co), 22, 0.1)                     # raw weights: masked entries are NOT zero here
    xt = torch.tensor(x.transpose(1, 4, 0, 2, 3), dtype=torch.float64, requires_grad=True)          # N, C, D, H, W
    wt = torch.tensor((w * mask[..., None, None]).transpose(4, 3, 0, 1, 2), dtype=torch.float64, requires_grad=True)    # Co, Ci, 2, 3, 3
    y = F.conv3d(xt, wt)
    dy = _rnd(tuple(y.shape), 23, gs)
    (y * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    wp = np.zeros((2, 3, 3, c4(Ci), c4(Co)), np.float32)
    wp[..., :Ci, :Co] = w * mask[..., None, None]
    pf, pd = nn.TcPlan.get('pc', False, Ci, Co), nn.TcPlan.get('pc', True, Ci, Co)
    assert pf is not None and pd is not None
    yg = pf.run(_cuda(x), _cuda(wp)).cpu().numpy()                                                   # D-1, N, H-2, W-2, c4(Co)
    ref = y.detach().numpy().transpose(2, 0, 3, 4, 1)
    assert yg.shape == ref.shape[:4] + (c4(Co),)
    err = _rel(yg[..., :Co], ref)
    print('tc plan pc fwd %d->%d: rel err %.2e' % (Ci, Co, err))
    assert err < 2e-5
    assert not yg[..., Co:].any()
    dyp = np.zeros(ref.shape[:4] + (c4(Co),), np.float32)
    dyp[..., :Co] = dy.transpose(2, 0, 3, 4, 1)
    dx = pd.run(_cuda(dyp), _cuda(wp)).cpu().numpy()
    refd = xt.grad.numpy().transpose(2, 0, 3, 4, 1)
    assert dx.shape == refd.shape
    err = _rel(dx, refd)
    print('tc plan pc dgrad: rel err %.2e (gradient scale %g)' % (err, gs))
    assert err < 2e-5
    # filter gradient; the masked taps get exact zeros (code/probclass.py:252-253: the mask multiplies the variable)
    pw = nn.TcWgradPlan.get('pc', Ci, Co)
    assert pw is not None
    dw = pw.run(_cuda(x), _cuda(dyp), torch.full((2, 3, 3, c4(Ci), c4(Co)), 7.0, device='cuda')).cpu().numpy()
    refw = wt.grad.numpy().transpose(2, 3, 4, 1, 0) * mask[..., None, None]
    err = _rel(dw[..., :Ci, :Co], refw)
    print('tc plan pc wgrad: rel err %.2e' % err)
    assert err < 2e-5
    assert not dw[1, 2].any() and not dw[1, 1, 2].any() and not dw[..., Ci:, :].any() and not dw[..., Co:].any()
    # unmasked raw weights give the same result: the plan never reads the masked taps
    wraw = np.zeros_like(wp)
    wraw[..., :Ci, :Co] = w
    assert np.array_equal(pf.run(_cuda(x), _cuda(wraw)).cpu().numpy(), yg) or _rel(pf.run(_cuda(x), _cuda(wraw)).cpu().numpy(), yg) < 1e-6


def test_tc_plan_unsupported_shapes_fall_back():
    from imgcomp_cvpr_b200 import nn
    assert nn.TcPlan.get('conv5s2', False, 3, 64) is None          # h1
    assert nn.TcPlan.get('conv5s2', True, 128, 33) is None         # to_bn's data gradient (33 input channels)
    assert nn.TcPlan.get('pc', False, 1, 24) is None               # first context-model layer (one input channel)


def test_tc_plan_pack_map_covers_every_weight_once():
    """the pack map of a plan (derived from the host packers through base-2047 digits) sends every real weight of the op to
    exactly one hi element of the packed image; everything else is zero padding"""
    from imgcomp_cvpr_b200 import _lib, nn
    for kind, data_grad, cin, cout, n_w in (('conv5s2', False, 64, 128, 25 * 64 * 128), ('conv5s2', True, 64, 128, 25 * 64 * 128),
                                            ('tconv5s2', False, 128, 64, 25 * 128 * 64), ('tconv5s2', True, 128, 64, 25 * 128 * 64),
                                            ('tconv5s2', False, 64, 3, 25 * 64 * 3), ('pc', False, 24, 24, 14 * 24 * 24),
                                            ('pc', True, 24, 6, 14 * 24 * 6)):
        plan = nn.TcPlan.get(kind, data_grad, cin, cout)
        L = _lib.lib()
        n = L.ic_nn_tc_plan_map(plan.h, None, 0)
        m = np.empty(n, np.int32)
        assert L.ic_nn_tc_plan_map(plan.h, m.ctypes.data, n) == n
        used = m[m >= 0]
        assert used.size == n_w and np.unique(used).size == n_w, (kind, data_grad, used.size, n_w)
        c4 = lambda v: (v + 3) // 4 * 4
        taps = 18 if kind == 'pc' else 25
        assert used.max() < taps * c4(cin) * c4(cout)
