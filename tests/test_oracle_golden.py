"""The oracle (oracle/imgcomp_oracle.py, numpy) against the golden vectors that
were produced by running the reference's own modules on the TF1 shim
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, symbol_margin
from oracle import imgcomp_oracle as O


@pytest.mark.parametrize('name', sorted(GOLDEN_CASES))
def test_val_graph_matches_reference_run(name, synth):
    ae_name, _ = GOLDEN_CASES[name]
    ae_cfg, pc_cfg, W = synth(ae_name)
    g = load_golden(name)
    r = O.val_forward(g['x_u8'], W, ae_cfg.num_chan_bn)
    r64 = O.encode(g['x_u8'].astype(np.float64), W, ae_cfg.num_chan_bn, dtype=np.float64)
    enc = r['enc']
    np.testing.assert_allclose(enc['z'], g['z'], atol=2e-4, rtol=0)
    # symbols: identical wherever the float64 latent is not within eps of a decision boundary
    margin = symbol_margin(r64['z'], W['autoencoder/encoder/centers'])
    safe = margin > 1e-4
    assert (enc['symbols'][safe] == g['symbols'][safe]).all()
    assert (enc['symbols'] != g['symbols']).mean() < 1e-3
    same = enc['symbols'] == g['symbols']
    np.testing.assert_allclose(enc['qbar'][same], g['qbar'][same], atol=1e-6)
    np.testing.assert_allclose(r['bitcost'][same], g['bitcost'][same], atol=2e-3)
    np.testing.assert_allclose(r['bpp'], g['bpp'], atol=1e-4)
    np.testing.assert_allclose(r['ms_ssim'], g['ms_ssim_np'], atol=1e-4)
    if 'x_out' in g:
        np.testing.assert_allclose(r['x_out'], g['x_out'], atol=5e-3)
        assert (r['x_out_u8'] != g['x_out_u8']).mean() < 1e-3
    if 'ms_ssim_tf_raises' in g:
        with pytest.raises((RuntimeError, ValueError)):
            O.ms_ssim_tf(g['x_u8'].astype(np.float32), r['x_out'])
    else:
        v = O.ms_ssim_tf(g['x_u8'].astype(np.float32), r['x_out'])[0]
        np.testing.assert_allclose(v, g['ms_ssim_tf'], atol=1e-4)


def test_real_bpp_frequencies_match_reference_loop(synth):
    """One batched pass == the reference's per-symbol PredictionNetwork loop
    (probclass.py:441-476 driven by bit_counter.py:85-134)."""
    ae_cfg, pc_cfg, W = synth('cvpr/low')
    g = load_golden('tiny_low_1x64x64')
    syms = g['symbols'][0].astype(np.int64)
    centers = W['autoencoder/encoder/centers']
    f = O.pc_freqs_volume(syms, W, centers)
    assert f.shape == g['freqs'].shape
    # float32 softmax * 1e9 truncated: allow the last float32 ulp (~60 counts at p~1)
    assert np.abs(f - g['freqs']).max() <= 128
    # literal per-context evaluation agrees with the batched pass
    sp = O.pad_for_probclass3d(syms, 9, 0)
    for (c, y, x) in ((0, 0, 0), (3, 2, 5), (31, 7, 7), (17, 0, 7)):
        fc = O.pc_freqs_context(sp[c:c + 5, y:y + 9, x:x + 9], W, centers)
        assert np.abs(fc - f[c, y, x]).max() <= 128
    bits_theory = -np.log2(f / f.sum(-1, keepdims=True))
    picked = np.take_along_axis(bits_theory, syms[..., None], -1).sum()
    assert abs(picked - float(g['theory_bits'])) < 1.0
    assert abs(int(g['real_bits']) - picked) < 50          # bit_counter.py:51


def test_msssim_pairs_match_reference_functions():
    g = load_golden('msssim_pairs')
    for tag in 'abc':
        x, y = g['x_' + tag], g['y_' + tag]
        for i in range(x.shape[0]):
            v = O.ms_ssim_np(x[i:i + 1].transpose(0, 2, 3, 1), y[i:i + 1].transpose(0, 2, 3, 1))[0]
            np.testing.assert_allclose(v, g['np_' + tag][i], atol=1e-9)
        v = O.ms_ssim_tf(x.astype(np.float32), y.astype(np.float32))[0]
        np.testing.assert_allclose(v, g['tf_' + tag], atol=2e-6)
        v64 = O.ms_ssim_tf(x.astype(np.float64), y.astype(np.float64), dtype=np.float64)[0]
        np.testing.assert_allclose(v64, g['tf_' + tag], atol=2e-5)


def test_float64_truth_agrees_with_float32_oracle(synth):
    ae_cfg, pc_cfg, W = synth('cvpr/low')
    g = load_golden('cfg1_low_1x128x128')
    r32 = O.val_forward(g['x_u8'], W, 32)
    r64 = O.val_forward(g['x_u8'], W, 32, dtype=np.float64)
    margin = symbol_margin(r64['enc']['z'], W['autoencoder/encoder/centers'])
    safe = margin > 1e-4
    assert (r32['enc']['symbols'][safe] == r64['enc']['symbols'][safe]).all()
    np.testing.assert_allclose(r32['bpp'], r64['bpp'], atol=1e-4)
    np.testing.assert_allclose(r32['ms_ssim'], r64['ms_ssim'], atol=1e-4)


def test_training_oracle_in_inference_mode_matches_golden(synth):
    """oracle/train_oracle.py (torch autograd restatement of the training graph), with batch norm put in
    inference mode, must reproduce the reference-run val graph: same symbols, bit cost, reconstruction."""
    import torch
    from oracle import train_oracle as T
    ae_cfg, pc_cfg, W = synth('cvpr/low')
    g = load_golden('tiny_low_1x64x64')
    P = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in W.items()}
    x = torch.tensor(g['x_u8'].astype(np.float64))
    enc = T.encode(x, P, ae_cfg.arch_param_B, False, {})
    x_out = T.decode(enc['qhard'], P, ae_cfg.arch_param_B, False, {})
    bc, _ = T.pc_bitcost(enc['qbar'], enc['symbols'], P, float(W['autoencoder/encoder/centers'][0]))
    assert np.array_equal(enc['symbols'].numpy(), g['symbols'])
    np.testing.assert_allclose(enc['z'].numpy(), g['z'], atol=5e-5)
    np.testing.assert_allclose(bc.numpy(), g['bitcost'], atol=2e-4)
    np.testing.assert_allclose(x_out.numpy(), g['x_out'], atol=2e-3)
    if 'ms_ssim_tf' in g:
        v = T.ms_ssim_tf(x, torch.tensor(g['x_out'].astype(np.float64)))
        np.testing.assert_allclose(float(v), float(g['ms_ssim_tf']), rtol=2e-5)


def test_training_oracle_gradients_are_consistent(synth):
    """finite differences on a few parameters of a tiny training step (float64)"""
    import copy
    from oracle import train_oracle as T
    from imgcomp_cvpr_b200 import weights as wm
    ae_cfg, pc_cfg, W = synth('cvpr/low')
    x = wm.synthetic_images(2, 32, 32, seed=11).astype(np.float64)
    r = T.training_step(x, W, ae_cfg, pc_cfg)
    assert np.isfinite(r['total_loss'])
    rng = np.random.RandomState(0)
    for name in ['autoencoder/decoder/h13/weights', 'autoencoder/decoder/h13/BatchNorm/beta',
                 'probclass3d/logits/conv3d_conv2_mask/biases']:
        g = r['grads'][name]
        idx = tuple(rng.randint(0, s) for s in g.shape)
        eps = 1e-6
        Wp, Wm = copy.copy(W), copy.copy(W)
        Wp[name] = W[name].astype(np.float64).copy(); Wp[name][idx] += eps
        Wm[name] = W[name].astype(np.float64).copy(); Wm[name][idx] -= eps
        fd = (T.training_step(x, Wp, ae_cfg, pc_cfg)['total_loss'] - T.training_step(x, Wm, ae_cfg, pc_cfg)['total_loss']) / (2 * eps)
        assert abs(fd - g[idx]) <= 2e-4 * max(1.0, abs(fd)), (name, fd, g[idx])


# ----------------------------------------------------------------------------
# training step: oracle/train_oracle.py against the reference's own graph code run on the autograd shim
# ----------------------------------------------------------------------------
def _proj(name, size):
    seed = int.from_bytes(name.encode(), 'little') % (2 ** 31 - 1)
    return np.random.RandomState(seed).standard_normal(size)


@pytest.mark.parametrize('tag', ['train_low_2x64x64', 'train_hi_2x80x48'])
def test_training_oracle_matches_reference_training_graph(tag, synth):
    """tests/golden/make_train_golden.py executed the UNMODIFIED reference modules (autoencoder, quantizer, probclass,
    ms_ssim, bits, train.get_loss / Distortions) with is_training=True on tests/tf1_shim/autograd.py and differentiated
    total_loss; the training oracle (what the CUDA step is tested against) must reproduce loss, batch statistics and
    the gradient of every one of the 219 variables (float64 both: 1e-6 relative, measured ~1e-12)."""
    import torch
    from imgcomp_cvpr_b200 import weights
    from oracle import train_oracle as T
    g = load_golden(tag)
    ae_name, N, H, W_, seed = (str(v) for v in g['meta'])
    ae_cfg, pc_cfg, Wt = synth(ae_name)
    x = weights.synthetic_images(int(N), int(H), int(W_), seed=int(seed))
    r = T.training_step(x, Wt, ae_cfg, pc_cfg, dtype=torch.float64, training=True)
    for k_ref, k in (('total_loss', 'total_loss'), ('H_real', 'H_real'), ('H_mask', 'H_mask'), ('pc_loss', 'pc_loss'),
                     ('d_loss_scaled', 'd_loss_scaled'), ('ms_ssim', 'ms_ssim')):
        ref = float(g['loss/' + k_ref])
        assert abs(r[k] - ref) <= 1e-9 * max(1.0, abs(ref)), (k, r[k], ref)
    assert np.array_equal(r['tensors']['symbols'].astype(np.uint8), g['symbols'])
    np.testing.assert_allclose(r['tensors']['bc'].sum(axis=(1, 2, 3)), g['bc_sum_per_image'], rtol=1e-10)
    np.testing.assert_allclose(r['tensors']['x_out'].mean(axis=(1, 2, 3)), g['x_out_mean_per_image'], rtol=1e-10)
    names = [str(n) for n in g['names']]
    assert sorted(r['grads']) == names
    worst = 0.0
    for i, name in enumerate(names):
        gr = np.asarray(r['grads'][name], np.float64)
        nref, pref = float(g['grad_norm'][i]), float(g['grad_proj'][i])
        assert abs(np.linalg.norm(gr) - nref) <= 1e-6 * max(nref, 1e-12), name
        pr = float(np.dot(gr.ravel(), _proj(name, gr.size)))
        assert abs(pr - pref) <= 1e-6 * max(nref * np.sqrt(gr.size), 1e-12), name
        worst = max(worst, abs(np.linalg.norm(gr) - nref) / max(nref, 1e-300))
    for key in list(g):
        if key.startswith('grad/'):
            ref = g[key]
            np.testing.assert_allclose(r['grads'][key[5:]], ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
        if key.startswith('bn_mean/'):
            s = key[len('bn_mean/'):]
            np.testing.assert_allclose(r['bn_stats'][s][0], g[key], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(r['bn_stats'][s][1], g['bn_var_unbiased/' + s], rtol=1e-9)
    print('%s: worst gradient-norm deviation %.1e over %d variables' % (tag, worst, len(names)))


@pytest.mark.parametrize('shape', [(2, 64, 64), (1, 70, 50), (3, 160, 160), (1, 47, 33)])
def test_msssim_gradient_oracle_matches_reference_module(shape):
    """value and gradient of oracle/train_oracle.ms_ssim_tf (what the CUDA MS-SSIM backward is tested against, odd sizes
    with REFLECT-padded levels included) against the reference's ms_ssim.MultiScaleSSIM differentiated on the autograd shim"""
    import torch
    from oracle import train_oracle as T
    g = load_golden('train_msssim_grad')
    N, H, W_ = shape
    rng = np.random.RandomState(5)
    a = rng.uniform(0, 255, size=(N, 3, H, W_)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 12, size=a.shape), 0, 255).astype(np.float32)
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    v = T.ms_ssim_tf(torch.tensor(a, dtype=torch.float64), bt)
    v.backward()
    gr = bt.grad.numpy()
    key = 'x'.join(str(s) for s in shape)
    assert abs(float(v.detach()) - float(g['value/' + key])) < 1e-12
    nref = float(g['grad_norm/' + key])
    assert abs(np.linalg.norm(gr) - nref) <= 1e-9 * nref
    assert abs(float(np.dot(gr.ravel(), _proj('msssim' + key, gr.size))) - float(g['grad_proj/' + key])) <= 1e-8 * nref * np.sqrt(gr.size)
    np.testing.assert_allclose(gr[0, :, :6, :6], g['grad_corner/' + key], rtol=1e-8, atol=1e-12 * nref)
    np.testing.assert_allclose(gr[-1, :, -6:, -6:], g['grad_tail/' + key], rtol=1e-8, atol=1e-12 * nref)


def test_golden_check_helper_on_the_oracle(synth):
    """conftest.check_training_against_golden (the helper the GPU training test uses) exercised on the CPU with the oracle's
    output: once with the regularisation terms in the gradients, once with them stripped and re-added by the helper;
    and it must reject a perturbed gradient."""
    import torch
    from conftest import check_training_against_golden
    from imgcomp_cvpr_b200 import weights
    from oracle import train_oracle as T
    g = load_golden('train_hi_2x80x48')
    ae_cfg, pc_cfg, Wt = synth('cvpr/hi')
    x = weights.synthetic_images(2, 80, 48, seed=21)
    r = T.training_step(x, Wt, ae_cfg, pc_cfg, dtype=torch.float64, training=True)
    assert check_training_against_golden(r, r['grads'], g, Wt, ae_cfg, False, 1e-9, 1e-6) < 1e-9
    stripped = {}
    for k, v in r['grads'].items():
        w = np.asarray(Wt[k], np.float64)
        if k.startswith('autoencoder/') and k.endswith('/weights'):
            v = v - ae_cfg.regularization_factor * w
        elif k.endswith('/centers'):
            v = v - ae_cfg.regularization_factor_centers * w
        stripped[k] = v
    assert check_training_against_golden(r, stripped, g, Wt, ae_cfg, True, 1e-9, 1e-6) < 1e-9
    bad = dict(r['grads'])
    k = 'autoencoder/decoder/h12/weights'
    bad[k] = bad[k] * 1.01
    with pytest.raises(AssertionError):
        check_training_against_golden(r, bad, g, Wt, ae_cfg, False, 1e-9, 1e-6)
