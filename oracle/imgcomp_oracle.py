"""CPU restatement (numpy) of the fab-jul/imgcomp-cvpr forward hot path.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Never imported by the product.

Every function cites the reference file:line (relative to /root/reference/code)
it follows.  The reference is TensorFlow-1.4 graph code; TF itself is not in
the tree, so the TF/slim kernel semantics (SAME padding, conv2d_transpose,
fused batch norm, softmax, conv3d VALID, REFLECT pad) are restated from their
published definitions.

PARITY STATUS: "parity unpinned" at the TF-kernel level -- the reference ships
no tests / golden vectors and TF-1.4 cannot be run here.  What IS pinned
(tests/test_oracle_golden.py, tests/golden/make_golden.py):
  * ms_ssim_np.MultiScaleSSIM           == reference function imported and run
  * arithmetic coder                    == reference module imported and run
  * gauss_kernel / _FSpecialGauss       == reference functions imported and run
  * graph structure of encode / decode / probclass / ms_ssim / quantizer /
    bit_counter == the reference modules executed on a numpy-backed TF1 shim
    (tests/tf1_shim) whose kernels are the restated semantics.

Layouts follow the reference: images / latents NCHW, conv2d weights HWIO,
conv2d_transpose weights [kh, kw, Cout, Cin], conv3d weights [D, H, W, in, out].
All arithmetic is done in ``dtype`` (float32 = "as reference", float64 = truth).
"""
import math

import numpy as np

# ----------------------------------------------------------------------------
# constants of the reference
# ----------------------------------------------------------------------------
# autoencoder.py:160-169
NORM_MEAN = np.array([121.85369873, 113.58860779, 100.63715363], dtype=np.float32)
NORM_VAR = np.array([4746.37695312, 4454.13964844, 4812.234375], dtype=np.float32)
BN_EPS = 1e-5                     # autoencoder.py:118
ARCH_N = 128                      # autoencoder.py:210
HARD_SIGMA = 1e7                  # quantizer.py:5
MSSSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)  # ms_ssim.py:165-166

_BACKEND = {'name': 'numpy'}


def set_backend(name):
    """'numpy' (authoritative) or 'torch' (same maths through torch-CPU conv
    kernels; used for the timed CPU baseline and for larger parity cases)."""
    assert name in ('numpy', 'torch')
    _BACKEND['name'] = name


# ----------------------------------------------------------------------------
# TF kernel semantics (not in the tree)
# ----------------------------------------------------------------------------
def same_pads(n, k, s):
    """TF 'SAME': out = ceil(n/s); pad_before = total // 2 (Appendix A.2)."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2, out


def _corr2d_valid(xp, w_hwio, stride):
    """VALID cross-correlation, xp NHWC (already padded), w [kh,kw,I,O]."""
    kh, kw, ci, co = w_hwio.shape
    if _BACKEND['name'] == 'torch':
        import torch
        import torch.nn.functional as F
        xt = torch.from_numpy(np.ascontiguousarray(xp.transpose(0, 3, 1, 2)))
        wt = torch.from_numpy(np.ascontiguousarray(w_hwio.transpose(3, 2, 0, 1)))
        y = F.conv2d(xt, wt, stride=stride)
        return y.numpy().transpose(0, 2, 3, 1)
    win = np.lib.stride_tricks.sliding_window_view(xp, (kh, kw), axis=(1, 2))
    win = win[:, ::stride, ::stride]                       # N,Ho,Wo,C,kh,kw
    wt = np.ascontiguousarray(w_hwio.transpose(2, 0, 1, 3))  # C,kh,kw,O
    return np.tensordot(win, wt, axes=3)


def conv2d_same(x, w, stride):
    """slim.conv2d core (autoencoder.py:222-237): NCHW in, HWIO weights, SAME,
    cross-correlation, no bias.  Returns NCHW."""
    kh, kw = w.shape[:2]
    pt, pb, _ = same_pads(x.shape[2], kh, stride)
    pl, pr, _ = same_pads(x.shape[3], kw, stride)
    xp = np.pad(x.transpose(0, 2, 3, 1), ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    y = _corr2d_valid(xp, w, stride)
    return np.ascontiguousarray(y.transpose(0, 3, 1, 2))


def conv2d_transpose_same(x, w, stride):
    """slim.conv2d_transpose core (autoencoder.py:251,264-265).  w is
    [kh,kw,Cout,Cin].  Defined as the gradient of the SAME/stride-s forward
    conv from s*n -> n: y[i] = sum_{o,t : s*o + t - pb = i} x[o] w[t]
    (Appendix A.2): full transposed conv cropped to [pb : pb + s*n]."""
    kh, kw, co, ci = w.shape
    n, _, h, ww = x.shape
    s = stride
    xu = np.zeros((n, (h - 1) * s + 1, (ww - 1) * s + 1, ci), dtype=x.dtype)
    xu[:, ::s, ::s, :] = x.transpose(0, 2, 3, 1)
    xp = np.pad(xu, ((0, 0), (kh - 1, kh - 1), (kw - 1, kw - 1), (0, 0)))
    wf = np.ascontiguousarray(w[::-1, ::-1].transpose(0, 1, 3, 2))  # kh,kw,Cin,Cout flipped
    full = _corr2d_valid(xp, wf, 1)                          # N,(h-1)s+kh,(w-1)s+kw,Cout
    pbh = same_pads(h * s, kh, s)[0]
    pbw = same_pads(ww * s, kw, s)[0]
    y = full[:, pbh:pbh + h * s, pbw:pbw + ww * s]
    return np.ascontiguousarray(y.transpose(0, 3, 1, 2))


def batch_norm_inference(x, bn, dtype):
    """slim.batch_norm, is_training=False (autoencoder.py:115-125, A.1):
    y = (x - mu) * rsqrt(var + eps) * gamma + beta, per channel (NCHW)."""
    g, b, mu, var = (bn[k].astype(dtype)[None, :, None, None]
                     for k in ('gamma', 'beta', 'moving_mean', 'moving_variance'))
    inv = (dtype(1.0) / np.sqrt(var + dtype(BN_EPS))).astype(dtype)
    return ((x - mu) * (inv * g) + b).astype(dtype)


# ----------------------------------------------------------------------------
# weights access (Appendix B variable schema)
# ----------------------------------------------------------------------------
def _conv_bn(W, scope, dtype):
    w = W[scope + '/weights'].astype(dtype)
    bn = {k: W[scope + '/BatchNorm/' + k] for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')}
    return w, bn


def _slim_conv(x, W, scope, stride, relu, dtype, transpose=False):
    """slim.conv2d / conv2d_transpose under _batch_norm_scope: conv -> BN -> act
    (autoencoder.py:106-113; no bias because normalizer_fn is set)."""
    w, bn = _conv_bn(W, scope, dtype)
    y = conv2d_transpose_same(x, w, stride) if transpose else conv2d_same(x, w, stride)
    y = batch_norm_inference(y.astype(dtype), bn, dtype)
    return np.maximum(y, dtype(0)) if relu else y


def _residual_block(x, W, scope, relu_first, dtype):
    """autoencoder.residual_block (autoencoder.py:274-287), num_conv2d=2:
    conv1 (+ReLU unless activation_fn=None was passed) -> conv2 (no act) -> + x."""
    y = _slim_conv(x, W, scope + '/conv1', 1, relu_first, dtype)
    y = _slim_conv(y, W, scope + '/conv2', 1, False, dtype)
    return y + x


def _res_stack(net, W, prefix, tag, final_scope, B, dtype):
    """The 5x3 residual blocks + final no-ReLU block + long skip shared by
    _CVPR._encode (autoencoder.py:224-234) and _decode (:252-262)."""
    r0 = net
    for b in range(B):
        rb = net
        for i in (1, 2, 3):
            net = _residual_block(net, W, '{}/res_block_{}_{}/{}_{}_{}'.format(prefix, tag, b, tag, b, i),
                                  True, dtype)
        net = net + rb
    net = _residual_block(net, W, prefix + '/' + final_scope, False, dtype)
    return net + r0


# ----------------------------------------------------------------------------
# quantizer.py
# ----------------------------------------------------------------------------
def _softmax_last(x):
    """tf.nn.softmax: exp(x - max) / sum (Eigen kernel)."""
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True, dtype=x.dtype)


def quantize(x, centers, sigma=1, dtype=np.float32):
    """quantizer.quantize/_quantize1d (quantizer.py:37-95), NCHW.
    Returns (qsoft, qhard, symbols int64)."""
    x = x.astype(dtype)
    c = centers.astype(dtype)
    dist = np.square(np.abs(x[..., None] - c))              # :75  (B,C,H,W,L)
    phi_soft = _softmax_last(dtype(-sigma) * dist)          # :79
    phi_hard = _softmax_last(dtype(-HARD_SIGMA) * dist)     # :82
    symbols = np.argmax(phi_hard, axis=-1).astype(np.int64)  # :84 first index on ties
    onehot = np.eye(len(c), dtype=dtype)[symbols]           # :85
    qsoft = _phi_times_centers(phi_soft, c)                 # :88
    qhard = _phi_times_centers(onehot, c)                   # :90
    return qsoft, qhard, symbols


def _phi_times_centers(phi, c):
    """quantizer.phi_times_centers (quantizer.py:98-100); summed in index order."""
    prod = phi * c
    acc = prod[..., 0].copy()
    for j in range(1, prod.shape[-1]):
        acc = acc + prod[..., j]
    return acc


# ----------------------------------------------------------------------------
# autoencoder.py
# ----------------------------------------------------------------------------
def normalize(x, dtype):
    """_Network._normalize, FIXED (autoencoder.py:136-144): divisor is
    np.sqrt(var + 1e-10) evaluated in float32 numpy, then broadcast."""
    div = np.sqrt(NORM_VAR + np.float32(1e-10))              # float32, as in the reference
    return ((x - NORM_MEAN.astype(dtype)[None, :, None, None]) / div.astype(dtype)[None, :, None, None]).astype(dtype)


def denormalize(x, dtype):
    """_Network._denormalize (autoencoder.py:146-154)."""
    mul = np.sqrt(NORM_VAR + np.float32(1e-10))
    return (x * mul.astype(dtype)[None, :, None, None] + NORM_MEAN.astype(dtype)[None, :, None, None]).astype(dtype)


def heatmap3d(bn, dtype):
    """_Network._get_heatmap3D (autoencoder.py:171-194)."""
    C = bn.shape[1] - 1
    hm2d = (dtype(1) / (dtype(1) + np.exp(-bn[:, 0]))) * dtype(C)           # sigmoid * C
    c = np.arange(C, dtype=dtype).reshape(1, C, 1, 1)
    return np.maximum(np.minimum(hm2d[:, None] - c, dtype(1)), dtype(0))


def encode(x, W, num_chan_bn=32, B=5, dtype=np.float32):
    """_Network.encode -> _CVPR._encode (autoencoder.py:50-58,218-244),
    is_training=False, heatmap=True, normalization=FIXED.
    x: float NCHW in [0,255].  Returns dict with the EncoderOutput fields plus
    'qsoft' and 'bn' (the to_bn output incl. heatmap channel)."""
    P = 'autoencoder/encoder'
    net = normalize(x.astype(dtype), dtype)
    net = _slim_conv(net, W, P + '/h1', 2, True, dtype)
    net = _slim_conv(net, W, P + '/h2', 2, True, dtype)
    net = _res_stack(net, W, P, 'enc', 'res_block_enc_final', B, dtype)
    bn = _slim_conv(net, W, P + '/to_bn', 2, False, dtype)
    assert bn.shape[1] == num_chan_bn + 1
    hm = heatmap3d(bn, dtype)
    z = hm * bn[:, 1:]                                          # _mask_with_heatmap :196-200
    centers = W[P + '/centers']
    qsoft, qhard, symbols = quantize(z, centers, 1, dtype)
    qbar = qsoft + (qhard - qsoft)                              # _quantize :133
    return dict(qbar=qbar, qhard=qhard, symbols=symbols, z=z, heatmap=hm, qsoft=qsoft, bn=bn)


def decode(q, W, B=5, dtype=np.float32):
    """_Network.decode -> _CVPR._decode (autoencoder.py:60-63,246-268)."""
    P = 'autoencoder/decoder'
    net = _slim_conv(q.astype(dtype), W, P + '/from_bn', 2, True, dtype, transpose=True)
    net = _res_stack(net, W, P, 'dec', 'dec_after_res', B, dtype)
    net = _slim_conv(net, W, P + '/h12', 2, True, dtype, transpose=True)
    net = _slim_conv(net, W, P + '/h13', 2, False, dtype, transpose=True)
    net = denormalize(net, dtype)
    return np.clip(net, dtype(0), dtype(255))                   # _clip_to_image_range :156-158


# ----------------------------------------------------------------------------
# probclass.py
# ----------------------------------------------------------------------------
def pc_masks(K=3):
    """create_first_mask / create_other_mask (probclass.py:150-176), DHW11."""
    shape = (K // 2 + 1, K, K)
    first = np.ones(shape, np.float32)
    first[-1, K // 2, K // 2:] = 0
    first[-1, K // 2 + 1:, :] = 0
    other = np.ones(shape, np.float32)
    other[-1, K // 2, K // 2 + 1:] = 0
    other[-1, K // 2 + 1:, :] = 0
    return first[..., None, None], other[..., None, None]


def pad_for_probclass3d(x, context_size=9, pad_value=0):
    """probclass.pad_for_probclass3d (probclass.py:268-292): pad depth front
    only, H/W both sides, constant."""
    pad = context_size // 2
    pads = [(0, 0)] * (x.ndim - 3) + [(pad, 0), (pad, pad), (pad, pad)]
    return np.pad(x, pads, mode='constant', constant_values=pad_value)


def _conv3d(x, W, scope, mask, relu, dtype):
    """probclass.conv3d (probclass.py:227-261): relu(conv3d_VALID(x, w*mask)+b)."""
    w = (W[scope + '/weights'].astype(dtype) * mask.astype(dtype))
    b = W[scope + '/biases'].astype(dtype)
    kd, kh, kw, ci, co = w.shape
    if _BACKEND['name'] == 'torch':
        import torch
        import torch.nn.functional as F
        xt = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 4, 1, 2, 3)))
        wt = torch.from_numpy(np.ascontiguousarray(w.transpose(4, 3, 0, 1, 2)))
        y = F.conv3d(xt, wt).numpy().transpose(0, 2, 3, 4, 1)
    else:
        win = np.lib.stride_tricks.sliding_window_view(x, (kd, kh, kw), axis=(1, 2, 3))  # N,D,H,W,C,kd,kh,kw
        y = np.tensordot(win, np.ascontiguousarray(w.transpose(3, 0, 1, 2, 4)), axes=4)
    y = (y + b).astype(dtype)
    return np.maximum(y, dtype(0)) if relu else y


def pc_logits(q_vol, W, dtype=np.float32):
    """_ResShallow._logits (probclass.py:214-221) on an NDHW1 volume (already
    padded, or a bare 5x9x9 context as in PredictionNetwork, probclass.py:441-442).
    Note the final conv keeps the default ReLU (probclass.py:220,233)."""
    P = 'probclass3d/logits'
    first, other = pc_masks(3)
    net = _conv3d(q_vol.astype(dtype), W, P + '/conv3d_conv0_mask', first, True, dtype)
    r = net
    net = _conv3d(net, W, P + '/res1/conv3d_conv1_mask', other, True, dtype)
    net = _conv3d(net, W, P + '/res1/conv3d_conv2_mask', other, False, dtype)
    net = net + r[:, 2:, 2:-2, 2:-2, :]                          # residual_block :196
    return _conv3d(net, W, P + '/conv3d_conv2_mask', other, True, dtype)


def pc_bitcost(q, symbols, W, pad_value, dtype=np.float32):
    """_Network3D.bitcost (probclass.py:63-106): pad -> logits ->
    softmax_cross_entropy_with_logits * log2(e).  q NCHW, symbols int NCHW."""
    qp = pad_for_probclass3d(q.astype(dtype), 9, dtype(pad_value))
    logits = pc_logits(qp[..., None], W, dtype)                  # N,C,h,w,L
    m = logits.max(axis=-1, keepdims=True)
    sh = logits - m
    lse = np.log(np.exp(sh).sum(axis=-1, dtype=dtype))
    picked = np.take_along_axis(sh, symbols[..., None], axis=-1)[..., 0]
    return ((lse - picked) * dtype(np.log2(np.e))).astype(dtype), logits


def bitcost_to_bpp(bitcost, input_batch_shape):
    """bits.bitcost_to_bpp (bits.py:4-20): sum(bits) / (prod(shape(x)) / 3)."""
    return bitcost.sum(dtype=bitcost.dtype) / bitcost.dtype.type(np.prod(input_batch_shape) / 3)


def freqs_from_logits(logits, resolution=1e9):
    """PredictionNetwork (probclass.py:443-444,465-476): softmax in float32,
    int64(trunc(pr * 1e9)), max(.,1)."""
    pr = _softmax_last(logits.astype(np.float32))
    f = (pr * np.float32(resolution)).astype(np.int64)
    return np.maximum(f, 1), pr


def pc_freqs_volume(symbols_chw, W, centers):
    """What the reference's per-symbol loop computes for every position
    (bit_counter.py:85-91 + probclass.py:441-476), evaluated in one pass:
    pad the *symbol* volume with symbol 0 (probclass.py:449-451), gather
    centres, run the logits network fully convolutionally.  -> int64 (C,h,w,L)."""
    sp = pad_for_probclass3d(symbols_chw, 9, 0)
    q = centers.astype(np.float32)[sp]
    logits = pc_logits(q[None, ..., None], W, np.float32)[0]
    return freqs_from_logits(logits)[0]


def pc_freqs_context(ctx_syms, W, centers):
    """PredictionNetwork.get_freqs on ONE (5,9,9) context -- the literal
    per-symbol evaluation of the reference (probclass.py:441-444,465-476)."""
    q = centers.astype(np.float32)[ctx_syms]
    logits = pc_logits(q[None, ..., None], W, np.float32)
    assert logits.shape[1:4] == (1, 1, 1)
    return freqs_from_logits(logits[0, 0, 0, 0])[0]


# ----------------------------------------------------------------------------
# ms_ssim.py  (TF float32 loss variant)
# ----------------------------------------------------------------------------
def gauss_kernel(sigma, size):
    """ms_ssim.gauss_kernel (ms_ssim.py:5-13): 2*(size//2)+1 taps, float64."""
    N = size // 2
    x = np.arange(-N, N + 1, 1.0)
    g = np.exp(-x * x / (2 * sigma * sigma))
    return g / np.sum(np.abs(g))


def _reflect_pad_hw(img, p1, p2):
    return np.pad(img, ((0, 0), (p1, p2), (p1, p2), (0, 0)), mode='reflect')


def _sep_valid(img, k):
    """Two tf.nn.conv2d VALID passes per channel: (1,K) then (K,1)
    (ms_ssim.py:31-43 / :52-63).  img NHWC."""
    K = len(k)
    win = np.lib.stride_tricks.sliding_window_view(img, K, axis=2)      # N,H,W',C,K
    h = np.zeros(win.shape[:-1], img.dtype)
    for t in range(K):
        h = h + win[..., t] * k[t]
    win = np.lib.stride_tricks.sliding_window_view(h, K, axis=1)        # N,H',W',C,K
    v = np.zeros(win.shape[:-1], img.dtype)
    for t in range(K):
        v = v + win[..., t] * k[t]
    return v


def gaussian_blur_tf(img, sigma, size, dtype):
    """ms_ssim.gaussian_blur (ms_ssim.py:16-43) incl. the operator-precedence
    quirk pad_w1 = total_pad + 1 // 2 = total_pad (:26); pads use the W size (:24)."""
    k = gauss_kernel(sigma, size).astype(dtype)
    total_pad = max(len(k) - img.shape[2], 0)
    p1, p2 = total_pad + 1 // 2, total_pad // 2
    return _sep_valid(_reflect_pad_hw(img, p1, p2), k)


def _ssim_for_multiscale_tf(a, b, dtype, max_val=255, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """ms_ssim._SSIMForMultiScale (ms_ssim.py:81-112), a/b NHWC."""
    _, H, Wd, _ = a.shape
    size = min(filter_size, H, Wd)
    sigma = size * filter_sigma / filter_size
    mu1 = gaussian_blur_tf(a, sigma, size, dtype)
    mu2 = gaussian_blur_tf(b, sigma, size, dtype)
    s11 = gaussian_blur_tf(a * a, sigma, size, dtype)
    s22 = gaussian_blur_tf(b * b, sigma, size, dtype)
    s12 = gaussian_blur_tf(a * b, sigma, size, dtype)
    mu11, mu22, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s11, s22, s12 = s11 - mu11, s22 - mu22, s12 - mu12
    c1 = dtype((k1 * max_val) ** 2)
    c2 = dtype((k2 * max_val) ** 2)
    v1 = dtype(2.0) * s12 + c2
    v2 = s11 + s22 + c2
    ssim = ((dtype(2.0) * mu12 + c1) * v1) / ((mu11 + mu22 + c1) * v2)
    return ssim.mean(dtype=dtype), (v1 / v2).mean(dtype=dtype)


def ms_ssim_tf(img1, img2, dtype=np.float32, data_format='NCHW'):
    """ms_ssim.MultiScaleSSIM (ms_ssim.py:115-186): ONE scalar for the batch.
    Returns (value, per-level ssim, per-level cs)."""
    if img1.shape != img2.shape:
        raise RuntimeError('Input images must have the same shape')
    if img1.ndim != 4:
        raise RuntimeError('Input images must have four dimensions')
    a, b = img1.astype(dtype), img2.astype(dtype)
    if data_format == 'NCHW':
        a, b = a.transpose(0, 2, 3, 1), b.transpose(0, 2, 3, 1)
    w = np.array(MSSSIM_WEIGHTS).astype(dtype)      # tf.convert_to_tensor(..., float32) :168
    box = np.array([0.5, 0.5], dtype)
    mssim, mcs = [], []
    for _ in range(len(w)):
        s, c = _ssim_for_multiscale_tf(a, b, dtype)
        mssim.append(s)
        mcs.append(c)
        # kernel_blur(pad=True): REFLECT pad (0,1), 2-tap separable, then ::2 (:46-64,179-181)
        a, b = (_sep_valid(_reflect_pad_hw(t, 0, 1), box)[:, ::2, ::2, :] for t in (a, b))
    mssim = np.array(mssim, dtype)
    mcs = np.array(mcs, dtype)
    val = np.prod(mcs[:-1] ** w[:-1], dtype=dtype) * (mssim[-1] ** w[-1])
    return dtype(val), mssim, mcs


# ----------------------------------------------------------------------------
# ms_ssim_np.py  (float64 metric on uint8, per image, used by val.py)
# ----------------------------------------------------------------------------
def fspecial_gauss_1d(size, sigma):
    """1-D factor of ms_ssim_np._FSpecialGauss (ms_ssim_np.py:113-124): the 2-D
    window exp(-(x^2+y^2)/2s^2)/sum is exactly outer(g,g) with g normalised."""
    radius = size // 2
    offset = 0.0
    start, stop = -radius, radius + 1
    if size % 2 == 0:
        offset = 0.5
        stop -= 1
    x = np.arange(offset + start, stop, 1.0)
    assert len(x) == size
    g = np.exp(-(x ** 2) / (2.0 * sigma ** 2))
    return g / g.sum()


def _box_reflect_ds(im):
    """scipy.ndimage.convolve(im, ones((1,2,2,1))/4, mode='reflect')[:, ::2, ::2]
    (ms_ssim_np.py:96,106-108).  For a size-2 kernel ndimage (centre index 1,
    flipped kernel) gives out[i] = (in[i] + in[i+1]) / 2 per axis -- checked
    against scipy -- with the 'reflect' (edge-duplicating) boundary in[n] := in[n-1];
    so even sizes get the plain 2x2 mean."""
    p = np.pad(im, ((0, 0), (0, 1), (0, 1), (0, 0)), mode='symmetric')
    s = (p[:, :-1, :-1] + p[:, 1:, :-1] + p[:, :-1, 1:] + p[:, 1:, 1:]) / 4.0
    return s[:, ::2, ::2]


def ms_ssim_np(img1, img2, max_val=255, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """ms_ssim_np.MultiScaleSSIM (ms_ssim_np.py:51-200) on NHWC uint8, float64.
    Separable evaluation of the reference's 2-D fftconvolve window."""
    if img1.shape != img2.shape:
        raise RuntimeError('Input images must have the same shape')
    if img1.ndim != 4:
        raise RuntimeError('Input images must have four dimensions')
    w = np.array(MSSSIM_WEIGHTS)
    a, b = img1.astype(np.float64), img2.astype(np.float64)
    mssim, mcs = [], []
    for _ in range(len(w)):
        _, H, Wd, _ = a.shape
        size = min(filter_size, H, Wd)
        sigma = size * filter_sigma / filter_size
        g = fspecial_gauss_1d(size, sigma)
        mu1, mu2 = _sep_valid(a, g), _sep_valid(b, g)
        s11 = _sep_valid(a * a, g) - mu1 * mu1
        s22 = _sep_valid(b * b, g) - mu2 * mu2
        s12 = _sep_valid(a * b, g) - mu1 * mu2
        c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
        v1 = 2.0 * s12 + c2
        v2 = s11 + s22 + c2
        mssim.append(np.mean(((2.0 * mu1 * mu2 + c1) * v1) / ((mu1 * mu1 + mu2 * mu2 + c1) * v2)))
        mcs.append(np.mean(v1 / v2))
        a, b = _box_reflect_ds(a), _box_reflect_ds(b)
    mssim, mcs = np.array(mssim), np.array(mcs)
    return float(np.prod(mcs[:-1] ** w[:-1]) * (mssim[-1] ** w[-1])), mssim, mcs


def psnr_u8(a, b):
    """val.psnr_np -> skimage compare_psnr on uint8 (val.py:227-237):
    10 log10(255^2 / mse), float64."""
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float(10 * np.log10(255.0 ** 2 / mse))


# ----------------------------------------------------------------------------
# val.py forward (cfg 1/2/4): the end-to-end restatement used for goldens
# ----------------------------------------------------------------------------
def val_forward(x_u8, W, num_chan_bn=32, dtype=np.float32):
    """val.validate graph, one batch (val.py:81-94): encode -> decode(qhard)
    -> bitcost(qbar, symbols, pad=centers[0]) -> bpp -> uint8 cast -> ms-ssim(np) -> psnr.
    x_u8: (N,3,H,W) uint8.  Metrics are per image as in the reference (N=1 there)."""
    x = x_u8.astype(dtype)
    enc = encode(x, W, num_chan_bn, dtype=dtype)
    x_out = decode(enc['qhard'], W, dtype=dtype)
    centers = W['autoencoder/encoder/centers']
    bc, logits = pc_bitcost(enc['qbar'], enc['symbols'], W, centers[0], dtype)
    x_out_u8 = x_out.astype(np.uint8)                        # tf.cast truncation, val.py:91
    per_img = []
    for i in range(x_u8.shape[0]):
        bpp = bitcost_to_bpp(bc[i:i + 1], x[i:i + 1].shape)
        ms = ms_ssim_np(x_u8[i:i + 1].transpose(0, 2, 3, 1), x_out_u8[i:i + 1].transpose(0, 2, 3, 1))[0]
        per_img.append((float(bpp), ms, psnr_u8(x_u8[i], x_out_u8[i])))
    return dict(enc=enc, x_out=x_out, x_out_u8=x_out_u8, bitcost=bc, logits=logits,
                bpp=np.array([p[0] for p in per_img]), ms_ssim=np.array([p[1] for p in per_img]),
                psnr=np.array([p[2] for p in per_img]))


# ----------------------------------------------------------------------------
# train.py loss (forward)
# ----------------------------------------------------------------------------
def get_loss(bc, heatmap, d_loss_scaled, H_target, beta, reg_enc=0.0, reg_dec=0.0, reg_pc=0.0):
    """train.get_loss (train.py:303-336), float32."""
    H_real = np.float32(bc.mean(dtype=np.float32))
    H_mask = np.float32((bc * heatmap).mean(dtype=np.float32)) if heatmap is not None else H_real
    H_soft = np.float32(0.5) * (H_mask + H_real)
    pc_loss = np.float32(beta) * np.maximum(H_soft - np.float32(H_target), np.float32(0))
    total = np.float32(d_loss_scaled) + pc_loss + np.float32(reg_pc + reg_enc + reg_dec)
    return float(total), float(H_real), float(H_mask), float(pc_loss)


def mse_per_img(inp, otp, cast_to_int):
    """Distortions.get_mse_per_img (train.py:400-418)."""
    if cast_to_int:
        inp, otp = inp.astype(np.int32), otp.astype(np.int32)      # tf.cast truncates
    se = np.square(otp - inp).astype(np.float32)
    return se.mean(axis=(1, 2, 3), dtype=np.float32)


def psnr_per_img(inp, otp, cast_to_int):
    """Distortions.get_psnr_per_image (train.py:420-425)."""
    return (10 * np.log(255.0 * 255.0 / mse_per_img(inp, otp, cast_to_int)) / np.log(10.0)).astype(np.float32)
