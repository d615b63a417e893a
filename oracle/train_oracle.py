"""CPU oracle of ONE TRAINING STEP of the reference (TEST INFRASTRUCTURE ONLY:
imported by tests/ and bench.py's cpu_baseline leg, never by the product).

Restates, with torch autograd on the CPU doing the differentiation, the graph
code/train.py:86-132 builds for is_training=True:

    enc = ae.encode(x, True)                      autoencoder.py:50-58,218-244 (batch-statistics BN, A.1)
    x_out = ae.decode(enc.qbar, True)             :60-63,246-268
    bc = pc.bitcost(stop_gradient(enc.qbar), enc.symbols, True, pad_value=centers[0])   train.py:103-105
    d = Distortions(config, x, x_out, True)       train.py:352-431
    total_loss = d.d_loss_scaled + beta * max(0.5 * (mean(bc * heatmap) + mean(bc)) - H_target, 0) + reg   :303-336
    two Adam optimisers (ae vars / probclass vars)  train.py:339-349, training_helpers.py:22-48

The forward arithmetic follows oracle/imgcomp_oracle.py (pinned against the reference-run goldens in inference mode:
tests/test_oracle_golden.py checks that this module, put in inference mode, reproduces it).
Training mode is pinned at GRAPH LEVEL too: tests/golden/make_train_golden.py executes the UNMODIFIED reference modules
(autoencoder, quantizer, probclass, ms_ssim, bits, and train.py's get_loss / Distortions) with is_training=True on
tests/tf1_shim/autograd.py (a torch-autograd stand-in for the TF-1.4 symbols) and differentiates total_loss;
tests/test_oracle_golden.py::test_training_oracle_matches_reference_training_graph holds this module to those vectors
(loss components, batch statistics, the gradient of all 219 variables: agreement ~1e-14 in float64).  What stays
unpinned is the kernel level, as for inference: no TF binary ever produced a number here (TF 1.4 is not installable).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import imgcomp_oracle as O

BN_EPS = 1e-5
BN_DECAY = 0.9


def _pads(n, k, s):
    return O.same_pads(n, k, s)[:2]


def conv2d_same(x, w, stride):
    """x NCHW, w HWIO (TF), SAME"""
    kh, kw = w.shape[:2]
    pt, pb = _pads(x.shape[2], kh, stride)
    pl, pr = _pads(x.shape[3], kw, stride)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w.permute(3, 2, 0, 1), stride=stride)


def conv2d_transpose_same(x, w, stride):
    """x NCHW, w [kh,kw,Cout,Cin] (TF conv2d_transpose): gradient of the SAME conv 2n -> n (A.2)"""
    kh, kw = w.shape[:2]
    full = F.conv_transpose2d(x, w.permute(3, 2, 0, 1), stride=stride)
    n_h, n_w = x.shape[2] * stride, x.shape[3] * stride
    pb_h = _pads(n_h, kh, stride)[0]
    pb_w = _pads(n_w, kw, stride)[0]
    return full[:, :, pb_h:pb_h + n_h, pb_w:pb_w + n_w]


def batch_norm(x, P, scope, training, stats):
    g, b = P[scope + '/BatchNorm/gamma'], P[scope + '/BatchNorm/beta']
    if training:
        mu = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.numel() // x.shape[1]
        stats[scope] = (mu.detach(), (var * n / max(n - 1, 1)).detach())     # what feeds the moving averages
    else:
        mu, var = P[scope + '/BatchNorm/moving_mean'], P[scope + '/BatchNorm/moving_variance']
    inv = 1.0 / torch.sqrt(var + BN_EPS)
    return (x - mu[None, :, None, None]) * (inv * g)[None, :, None, None] + b[None, :, None, None]


def slim_conv(x, P, scope, stride, relu, training, stats, transpose=False):
    w = P[scope + '/weights']
    y = conv2d_transpose_same(x, w, stride) if transpose else conv2d_same(x, w, stride)
    y = batch_norm(y, P, scope, training, stats)
    return torch.relu(y) if relu else y


def res_stack(net, P, prefix, tag, final_scope, B, training, stats):
    def block(x, scope, relu_first):
        y = slim_conv(x, P, scope + '/conv1', 1, relu_first, training, stats)
        y = slim_conv(y, P, scope + '/conv2', 1, False, training, stats)
        return y + x
    r0 = net
    for b in range(B):
        rb = net
        for i in (1, 2, 3):
            net = block(net, '{}/res_block_{}_{}/{}_{}_{}'.format(prefix, tag, b, tag, b, i), True)
        net = net + rb
    net = block(net, prefix + '/' + final_scope, False)
    return net + r0


def encode(x, P, B, training, stats):
    dt = x.dtype
    mean = torch.tensor(O.NORM_MEAN, dtype=dt)[None, :, None, None]
    div = torch.tensor(np.sqrt(O.NORM_VAR + np.float32(1e-10)), dtype=dt)[None, :, None, None]
    E = 'autoencoder/encoder'
    net = (x - mean) / div
    net = slim_conv(net, P, E + '/h1', 2, True, training, stats)
    net = slim_conv(net, P, E + '/h2', 2, True, training, stats)
    net = res_stack(net, P, E, 'enc', 'res_block_enc_final', B, training, stats)
    bn = slim_conv(net, P, E + '/to_bn', 2, False, training, stats)
    C = bn.shape[1] - 1
    hm2d = torch.sigmoid(bn[:, 0]) * C
    c = torch.arange(C, dtype=dt).reshape(1, C, 1, 1)
    hm = torch.clamp(hm2d[:, None] - c, 0, 1)                   # minimum(maximum(., 0), 1)
    z = hm * bn[:, 1:]
    centers = P[E + '/centers']
    dist = (z[..., None] - centers) ** 2
    qsoft = (torch.softmax(-dist, -1) * centers).sum(-1)
    d = dist.detach()
    symbols = (d == d.min(-1, keepdim=True).values).to(torch.uint8).argmax(-1)    # first index of the minimum
    qhard = centers.detach()[symbols]
    qbar = qsoft + (qhard - qsoft).detach()
    return dict(qbar=qbar, qhard=qhard, symbols=symbols, z=z, heatmap=hm, qsoft=qsoft, bn=bn)


def decode(q, P, B, training, stats):
    D = 'autoencoder/decoder'
    dt = q.dtype
    net = slim_conv(q, P, D + '/from_bn', 2, True, training, stats, transpose=True)
    net = res_stack(net, P, D, 'dec', 'dec_after_res', B, training, stats)
    net = slim_conv(net, P, D + '/h12', 2, True, training, stats, transpose=True)
    net = slim_conv(net, P, D + '/h13', 2, False, training, stats, transpose=True)
    mean = torch.tensor(O.NORM_MEAN, dtype=dt)[None, :, None, None]
    mul = torch.tensor(np.sqrt(O.NORM_VAR + np.float32(1e-10)), dtype=dt)[None, :, None, None]
    return torch.clamp(net * mul + mean, 0, 255)


def pc_bitcost(q, symbols, P, pad_value):
    """q NCHW (treated as a constant by the caller), symbols int64 NCHW -> bits NCHW, logits N,C,h,w,L"""
    first, other = (torch.tensor(m[..., 0, 0], dtype=q.dtype) for m in O.pc_masks(3))
    x = F.pad(q, (4, 4, 4, 4, 4, 0), value=float(pad_value))[:, None]            # N,1,D,H,W
    S = 'probclass3d/logits'

    def conv3d(t, scope, mask, relu):
        w = P[scope + '/weights'] * mask[:, :, :, None, None]                       # DHWio
        y = F.conv3d(t, w.permute(4, 3, 0, 1, 2)) + P[scope + '/biases'][None, :, None, None, None]
        return torch.relu(y) if relu else y
    net = conv3d(x, S + '/conv3d_conv0_mask', first, True)
    r = net
    net = conv3d(net, S + '/res1/conv3d_conv1_mask', other, True)
    net = conv3d(net, S + '/res1/conv3d_conv2_mask', other, False)
    net = net + r[:, :, 2:, 2:-2, 2:-2]
    logits = conv3d(net, S + '/conv3d_conv2_mask', other, True).permute(0, 2, 3, 4, 1)    # N,C,h,w,L
    bc = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), symbols.reshape(-1), reduction='none')
    return bc.reshape(symbols.shape) * math.log2(math.e), logits


def _gauss(sigma, size, dt):
    return torch.tensor(O.gauss_kernel(sigma, size), dtype=dt)


def _sep_valid(img, k):
    """img NCHW, per-channel (1,K) then (K,1) VALID"""
    C = img.shape[1]
    kw = k.reshape(1, 1, 1, -1).repeat(C, 1, 1, 1)
    kh = k.reshape(1, 1, -1, 1).repeat(C, 1, 1, 1)
    return F.conv2d(F.conv2d(img, kw, groups=C), kh, groups=C)


def ms_ssim_tf(a, b):
    """ms_ssim.MultiScaleSSIM (code/ms_ssim.py:115-186), NCHW, one scalar for the batch"""
    dt = a.dtype
    w = torch.tensor(np.array(O.MSSSIM_WEIGHTS), dtype=dt)         # tf.convert_to_tensor(weights, float32) (:168): rounded to `dt`
    box = torch.tensor([0.5, 0.5], dtype=dt)
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    mssim, mcs = [], []
    for _ in range(len(w)):
        H, Wd = a.shape[2], a.shape[3]
        size = min(11, H, Wd)
        k = _gauss(size * 1.5 / 11, size, dt)
        # ms_ssim.gaussian_blur (code/ms_ssim.py:16-43): REFLECT pad (total_pad + 1 // 2, total_pad // 2) on H and W,
        # total_pad taken from the W size (the operator-precedence quirk of :26 included)
        total_pad = max(len(k) - Wd, 0)
        p1, p2 = total_pad + 1 // 2, total_pad // 2

        def blur(t, k=k, p1=p1, p2=p2):
            if p1 or p2:
                t = F.pad(t, (p1, p2, p1, p2), mode='reflect')
            return _sep_valid(t, k)
        mu1, mu2 = blur(a), blur(b)
        s11, s22, s12 = blur(a * a) - mu1 * mu1, blur(b * b) - mu2 * mu2, blur(a * b) - mu1 * mu2
        v1, v2 = 2.0 * s12 + c2, s11 + s22 + c2
        mssim.append((((2.0 * mu1 * mu2 + c1) * v1) / ((mu1 * mu1 + mu2 * mu2 + c1) * v2)).mean())
        mcs.append((v1 / v2).mean())
        a, b = (_sep_valid(F.pad(t, (0, 1, 0, 1), mode='reflect'), box)[:, :, ::2, ::2] for t in (a, b))
    val = mssim[-1] ** w[-1]
    for l in range(len(w) - 1):
        val = val * mcs[l] ** w[l]
    return val


def l2(t):
    return (t ** 2).sum() / 2


def training_step(x, W, ae_cfg, pc_cfg, dtype=torch.float64, training=True):
    """x NCHW float/uint8 in [0,255]; W dict TF name -> ndarray.
    -> dict(loss components (floats), grads {name: ndarray}, bn_stats {scope: (mean, unbiased var)},
            tensors {bc, heatmap, x_out, symbols, qbar})."""
    P = {}
    for k, v in W.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        trainable = not (k.endswith('moving_mean') or k.endswith('moving_variance'))
        P[k] = t.requires_grad_(trainable)
    xt = torch.tensor(np.asarray(x), dtype=dtype)
    stats = {}
    enc = encode(xt, P, ae_cfg.arch_param_B, training, stats)
    x_out = decode(enc['qbar'], P, ae_cfg.arch_param_B, training, stats)
    centers = P['autoencoder/encoder/centers']
    bc, logits = pc_bitcost(enc['qbar'].detach(), enc['symbols'], P, centers[0].item())
    if ae_cfg.distortion_to_minimize == 'ms_ssim':
        msssim = ms_ssim_tf(xt, x_out)
        d_loss = ae_cfg.K_ms_ssim * (1 - msssim)
    elif ae_cfg.distortion_to_minimize == 'mse':
        msssim = None
        d_loss = ((x_out - xt) ** 2).mean(dim=(1, 2, 3)).mean()
    else:
        msssim = None
        mse = ((x_out - xt) ** 2).mean(dim=(1, 2, 3))
        d_loss = ae_cfg.K_psnr - (10 * torch.log10(255.0 * 255.0 / mse)).mean()
    H_real = bc.mean()
    H_mask = (bc * enc['heatmap']).mean()
    H_soft = 0.5 * (H_mask + H_real)
    pc_loss = ae_cfg.beta * torch.clamp(H_soft - ae_cfg.H_target, min=0)
    f = ae_cfg.regularization_factor
    reg_enc = sum(f * l2(v) for k, v in P.items() if k.startswith('autoencoder/encoder/') and k.endswith('/weights'))
    if ae_cfg.regularization_factor_centers != 0:
        reg_enc = reg_enc + ae_cfg.regularization_factor_centers * l2(centers)
    reg_dec = sum(f * l2(v) for k, v in P.items() if k.startswith('autoencoder/decoder/') and k.endswith('/weights'))
    reg_pc = 0.0
    if pc_cfg.regularization_factor is not None:
        reg_pc = pc_cfg.regularization_factor * sum(l2(v) for k, v in P.items()
                                                    if k.startswith('probclass3d/') and k.endswith('/weights'))
    total = d_loss + pc_loss + reg_enc + reg_dec + reg_pc
    total.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in P.items() if v.requires_grad}
    total, d_loss, pc_loss, H_real, H_mask = (t.detach() for t in (total, d_loss, pc_loss, H_real, H_mask))
    return dict(total_loss=float(total), d_loss_scaled=float(d_loss), pc_loss=float(pc_loss), H_real=float(H_real),
                H_mask=float(H_mask), ms_ssim=None if msssim is None else float(msssim.detach()),
                reg=float((reg_enc + reg_dec + reg_pc).detach()), grads=grads,
                bn_stats={k: (m.numpy(), v.numpy()) for k, (m, v) in stats.items()},
                tensors=dict(bc=bc.detach().numpy(), heatmap=enc['heatmap'].detach().numpy(),
                             x_out=x_out.detach().numpy(), symbols=enc['symbols'].numpy(),
                             qbar=enc['qbar'].detach().numpy(), z=enc['z'].detach().numpy()))


def adam_update(w, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer._apply_dense: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); step counts from 1"""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    return w - lr_t * m / (np.sqrt(v) + eps), m, v


def learning_rate(cfg, global_step, num_itr_per_epoch):
    """training_helpers.create_learning_rate_tensor (code/training_helpers.py:22-35)"""
    if cfg.lr_schedule == 'FIXED':
        return cfg.lr_initial
    p = global_step / float(num_itr_per_epoch * cfg.lr_schedule_decay_interval)
    if cfg.lr_schedule_decay_staircase:
        p = math.floor(p)
    return cfg.lr_initial * cfg.lr_schedule_decay_rate ** p
