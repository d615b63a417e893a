"""Variable schema of the reference graph + seeded synthetic weights / images.

The reference's checkpoints are not in its tree (README.md:14-15), so every
measurement and parity test runs on synthetic weights.  Names and shapes follow
the TF variable scopes of the reference:
  autoencoder/encoder/{h1,h2,to_bn}/weights + /BatchNorm/{gamma,beta,moving_mean,moving_variance}
      (code/autoencoder.py:222-237), conv2d weights HWIO
  autoencoder/encoder/res_block_enc_{b}/enc_{b}_{i}/conv{j}/...      (autoencoder.py:227-231)
  autoencoder/encoder/res_block_enc_final/conv{j}/...                 (autoencoder.py:232)
  autoencoder/encoder/centers (L,)                                    (quantizer.py:11-15)
  autoencoder/decoder/{from_bn,h12,h13}/..., conv2d_transpose weights [kh,kw,Cout,Cin]
      (autoencoder.py:251,264-265), res_block_dec_{b}/dec_{b}_{i}/conv{j}, dec_after_res/conv{j}
  probclass3d/logits/conv3d_conv0_mask/{weights (2,3,3,1,k), biases}  (probclass.py:217,249-257)
  probclass3d/logits/res1/conv3d_conv{1,2}_mask/{weights (2,3,3,k,k), biases}
  probclass3d/logits/conv3d_conv2_mask/{weights (2,3,3,k,L), biases}
so a converted TF checkpoint (a dict name -> ndarray) can be loaded unchanged.
"""
import numpy as np

ARCH_N = 128          # code/autoencoder.py:210
BN_KEYS = ('gamma', 'beta', 'moving_mean', 'moving_variance')

# Fixed, deliberately UNSORTED centres in (-2, 2): the reference draws them
# U(-2,2) with a TF seed that cannot be reproduced (quantizer.py:28-31) and never
# sorts them; centres[0] is the probclass pad value (probclass.py:59-61).
DEFAULT_CENTERS = np.array([0.35, -1.70, 1.20, -0.60, 1.90, -1.15], dtype=np.float32)


def conv_scopes(B=5):
    """(scope, kind, kh, stride, cin, cout, relu) for every autoencoder conv,
    in execution order.  cout/cin of to_bn/from_bn are filled by the caller (C)."""
    enc, dec = [], []
    for b in range(B):
        for i in (1, 2, 3):
            for j in (1, 2):
                enc.append('autoencoder/encoder/res_block_enc_{0}/enc_{0}_{1}/conv{2}'.format(b, i, j))
                dec.append('autoencoder/decoder/res_block_dec_{0}/dec_{0}_{1}/conv{2}'.format(b, i, j))
    enc += ['autoencoder/encoder/res_block_enc_final/conv1', 'autoencoder/encoder/res_block_enc_final/conv2']
    dec += ['autoencoder/decoder/dec_after_res/conv1', 'autoencoder/decoder/dec_after_res/conv2']
    return enc, dec


def _xavier(rng, shape, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _bn(rng, W, scope, c, pre_std=1.0, gamma_nominal=1.0):
    """BN statistics scaled to the layer's nominal pre-BN std (see
    _synth_table.py) so that activations stay O(1) like in a trained net."""
    W[scope + '/BatchNorm/gamma'] = (gamma_nominal * rng.uniform(0.5, 1.5, size=c)).astype(np.float32)
    W[scope + '/BatchNorm/beta'] = (0.1 * gamma_nominal * rng.standard_normal(c)).astype(np.float32)
    W[scope + '/BatchNorm/moving_mean'] = (0.1 * pre_std * rng.standard_normal(c)).astype(np.float32)
    W[scope + '/BatchNorm/moving_variance'] = (pre_std ** 2 * rng.uniform(0.6, 1.4, size=c)).astype(np.float32)


def _conv(rng, W, scope, k, cin, cout, table, transpose=False, gamma_nominal=1.0):
    shape = (k, k, cout, cin) if transpose else (k, k, cin, cout)
    W[scope + '/weights'] = _xavier(rng, shape, k * k * cin, k * k * cout)
    _bn(rng, W, scope, cout, table.get(scope, 1.0), gamma_nominal)


def _res_convs(rng, W, scopes, n, table):
    """conv1: gamma ~ 1; conv2: gamma ~ 0.3 x trunk rms (= conv1's pre-BN std,
    xavier 128->128 has unit gain) so each branch adds ~30 % to the trunk."""
    for s in scopes:
        if s.endswith('conv1'):
            _conv(rng, W, s, 3, n, n, table)
        else:
            _conv(rng, W, s, 3, n, n, table, gamma_nominal=0.3 * table.get(s[:-1] + '1', 1.0))


def synthetic_weights(num_chan_bn=32, num_centers=6, pc_k=24, B=5, seed=0, bn_table=None):
    """Deterministic weights for ae_configs/cvpr/{low,med,hi} + pc res_shallow:
    xavier-uniform convs, non-trivial BN statistics, small non-zero pc biases.
    ``bn_table`` maps conv scope -> nominal pre-BN std (default: the committed
    table produced by tests/golden/calibrate_synth_bn.py); the residual trunk of
    this architecture doubles in amplitude every group (net = net + skip), so BN
    moving statistics must follow it or latents saturate."""
    if bn_table is None:
        from ._synth_table import SYNTH_PRE_BN_STD
        bn_table = SYNTH_PRE_BN_STD.get(num_chan_bn, {})
    rng = np.random.RandomState(seed)
    n, C = ARCH_N, num_chan_bn
    W = {}
    E, D = 'autoencoder/encoder', 'autoencoder/decoder'
    _conv(rng, W, E + '/h1', 5, 3, n // 2, bn_table)
    _conv(rng, W, E + '/h2', 5, n // 2, n, bn_table)
    enc, dec = conv_scopes(B)
    _res_convs(rng, W, enc, n, bn_table)
    _conv(rng, W, E + '/to_bn', 5, n, C + 1, bn_table)
    assert num_centers == len(DEFAULT_CENTERS)
    W[E + '/centers'] = DEFAULT_CENTERS.copy()
    _conv(rng, W, D + '/from_bn', 3, C, n, bn_table, transpose=True)
    _res_convs(rng, W, dec, n, bn_table)
    _conv(rng, W, D + '/h12', 5, n, n // 2, bn_table, transpose=True)
    _conv(rng, W, D + '/h13', 5, n // 2, 3, bn_table, transpose=True)
    P = 'probclass3d/logits'
    for scope, ci, co in ((P + '/conv3d_conv0_mask', 1, pc_k),
                          (P + '/res1/conv3d_conv1_mask', pc_k, pc_k),
                          (P + '/res1/conv3d_conv2_mask', pc_k, pc_k),
                          (P + '/conv3d_conv2_mask', pc_k, num_centers)):
        W[scope + '/weights'] = _xavier(rng, (2, 3, 3, ci, co), 18 * ci, 18 * co)
        W[scope + '/biases'] = (0.01 * rng.standard_normal(co)).astype(np.float32)
    return W


def synthetic_images(n, h, w, seed=1234, kind='smooth'):
    """Seeded uint8 NCHW images: 'smooth' = low-pass noise + gradients (natural-
    image-like, non-degenerate latents), 'noise' = iid randint(0,256)."""
    rng = np.random.RandomState(seed)
    if kind == 'noise':
        return rng.randint(0, 256, size=(n, 3, h, w)).astype(np.uint8)
    out = np.empty((n, 3, h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        img = np.zeros((3, h, w), np.float32)
        for scale, amp in ((32, 60.0), (8, 35.0), (2, 18.0), (1, 8.0)):
            gh, gw = -(-h // scale) + 1, -(-w // scale) + 1
            g = rng.standard_normal((3, gh, gw)).astype(np.float32)
            up = np.kron(g, np.ones((1, scale, scale), np.float32))[:, :h, :w]
            if scale > 1:      # cheap separable box blur to remove blockiness
                k = scale
                cs = np.cumsum(np.pad(up, ((0, 0), (k, 0), (0, 0)), mode='edge'), axis=1)
                up = (cs[:, k:] - cs[:, :-k]) / k
                cs = np.cumsum(np.pad(up, ((0, 0), (0, 0), (k, 0)), mode='edge'), axis=2)
                up = (cs[:, :, k:] - cs[:, :, :-k]) / k
            img += amp * up
        a, b, c = rng.uniform(-0.3, 0.3, 3)
        img += 128.0 + (a * (yy - h / 2) + b * (xx - w / 2))[None] + 20 * c
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out
