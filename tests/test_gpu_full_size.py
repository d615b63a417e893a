"""Oracle parity at the BASELINE.json sizes (run with -m gpu on the B200 box).

configs[1] (768x512, cvpr/low), configs[3] (512x512, cvpr/hi) and configs[2] (160x160 crops, cvpr/med, training
step) through the public host API in BOTH arithmetic modes against the CPU oracle run on the same seeded input --
what code/val.py:81-94 / code/train.py:86-132 compute.  The torch backend of the oracle needs ~1.5 s per image
here (float32 and float64 passes), so the full sizes are compared directly, not only through properties.

Gates (tolerances written where they are used):
  * symbols: identical to the float32 oracle wherever the float64 latent is further than EPS from a decision
    boundary (two float32 implementations of a 34-layer conv stack differ by summation order),
    and the TOTAL mismatch count of the tensor-core mode is bounded by what the float32 FFMA mode shows;
  * bpp and MS-SSIM within 1e-4 of the oracle (BASELINE.json north_star).
"""
import numpy as np
import pytest
import torch

from conftest import symbol_margin
from oracle import imgcomp_oracle as O
from oracle import train_oracle as T

pytestmark = pytest.mark.gpu

# name: (ae config, N, H, W, image seed)
CASES = {
    'cfg2_kodak_low_768x512': ('cvpr/low', 1, 768, 512, 1234),
    'cfg4_tile_hi_512x512': ('cvpr/hi', 1, 512, 512, 4321),
    'headline_low_512x512': ('cvpr/low', 1, 512, 512, 99),
}
# symbol margin: |z64 - decision boundary| above which a symbol must match the oracle
EPS = {'fp32': 1e-4, 'exact': 2e-4}
_oracle_cache = {}


def _oracle(name, W, C):
    """float32 val.py graph + float64 latent of the oracle, cached across the mode parametrisation"""
    if name not in _oracle_cache:
        from imgcomp_cvpr_b200 import weights as wm
        ae_name, N, H, Wd, seed = CASES[name]
        x = wm.synthetic_images(N, H, Wd, seed=seed)
        O.set_backend('torch')
        try:
            ref = O.val_forward(x, W, C)
            z64 = O.encode(x.astype(np.float64), W, C, dtype=np.float64)['z']
        finally:
            O.set_backend('numpy')
        _oracle_cache[name] = (x, ref, z64)
    return _oracle_cache[name]


@pytest.mark.parametrize('mode', ['fp32', 'exact'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_full_size_val_graph_against_oracle(name, mode, gpu_models):
    from imgcomp_cvpr_b200 import bits, ms_ssim_np
    ae_name = CASES[name][0]
    ae, pc, W = gpu_models(ae_name, mode)
    C = ae.config.num_chan_bn
    x, ref, z64 = _oracle(name, W, C)
    xc = torch.from_numpy(x).cuda()
    enc = ae.encode(xc, is_training=False)
    x_out = ae.decode(enc.qhard, is_training=False)
    bc = pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
    sym, z = enc.symbols.cpu().numpy(), enc.z.cpu().numpy()
    rsym = ref['enc']['symbols']
    centers = W['autoencoder/encoder/centers']
    margin = symbol_margin(z64, centers)
    mism = sym != rsym
    o_mism = int((rsym != O.quantize(z64, centers.astype(np.float64), 1, dtype=np.float64)[2]).sum())
    dz = np.abs(z - z64)
    print('%s/%s: symbols %d, mismatches vs float32 oracle %d (float32 oracle vs float64 oracle: %d), '
          'max|z-z64| %.2e mean %.2e, largest margin of a mismatch %.2e' % (
              name, mode, sym.size, mism.sum(), o_mism, dz.max(), dz.mean(), margin[mism].max() if mism.any() else 0.0))
    safe = margin > EPS[mode]
    assert (sym[safe] == rsym[safe]).all(), 'a symbol flipped %.1e away from a decision boundary' % margin[mism & safe].max()
    # the count that can legitimately differ: positions closer to a boundary than the float32 error of either side;
    # the float32 oracle itself differs from float64 on o_mism of them
    assert mism.sum() <= max(2 * o_mism, 8), (int(mism.sum()), o_mism)
    # float32 FFMA chains reach 3e-4 on cvpr/hi (65-column to_bn); the tensor core's truncating fp32 accumulate adds to that
    assert dz.max() < (4e-4 if mode == 'fp32' else 1e-3)
    bpp = bits.bitcost_to_bpp(bc, xc).item()
    print('   bpp %.6f oracle %.6f   ' % (bpp, ref['bpp'][0]))
    assert abs(bpp - ref['bpp'][0]) < 1e-4                                            # north star: 1e-4
    assert abs(pc.last_bits_per_image[0].item() / (x.shape[2] * x.shape[3]) - ref['bpp'][0]) < 1e-4
    # a flipped symbol changes the context of every position whose 5x9x9 causal window contains it: compare the rest
    near = np.zeros_like(mism)
    for n, c, y, x_ in np.argwhere(mism):
        near[n, c:c + 5, max(y - 4, 0):y + 5, max(x_ - 4, 0):x_ + 5] = True
    np.testing.assert_allclose(bc.cpu().numpy()[~near], ref['bitcost'][~near], atol=2e-3)
    ms = ms_ssim_np.MultiScaleSSIM_batch(xc, ae.extra['x_out_u8'], data_format='NCHW').cpu().numpy()
    print('   ms-ssim %.6f oracle %.6f' % (ms[0], ref['ms_ssim'][0]))
    np.testing.assert_allclose(ms, ref['ms_ssim'], atol=1e-4)                      # north star: 1e-4
    if not mism.any():
        np.testing.assert_allclose(x_out.cpu().numpy(), ref['x_out'], atol=2e-2)


@pytest.mark.parametrize('mode', ['fp32', 'exact'])
def test_cfg3_training_step_160_med_against_oracle(mode, synth):
    """BASELINE.json configs[2] shape (160x160 crops, cvpr/med; 2 images here -- batch statistics need > 1) in the
    strict float32 mode AND the shipped tensor-core mode against the float64 autograd oracle."""
    from imgcomp_cvpr_b200 import trainer, weights
    ae_cfg, pc_cfg, Wt = synth('cvpr/med')
    x = weights.synthetic_images(2, 160, 160, seed=160)
    tr = trainer.Trainer(ae_cfg, pc_cfg, Wt, num_itr_per_epoch=100, mode=mode)
    ref = T.training_step(x, Wt, ae_cfg, pc_cfg, dtype=torch.float64, training=True)
    out = tr.forward_backward(torch.from_numpy(x).cuda(), is_training=True, update_moving=False)
    sym = out['tensors']['symbols'].cpu().numpy()
    mism = sym != ref['tensors']['symbols']
    z64 = ref['tensors']['z'] if 'z' in ref['tensors'] else None
    print('cfg3/%s: symbol mismatches %d / %d' % (mode, mism.sum(), sym.size))
    if z64 is not None and mism.any():
        m = symbol_margin(np.asarray(z64, np.float64), Wt['autoencoder/encoder/centers'])
        assert m[mism].max() < EPS[mode], m[mism].max()
    for k in ('total_loss', 'd_loss_scaled', 'pc_loss', 'H_real', 'H_mask', 'ms_ssim', 'reg'):
        print('  %-14s gpu %.6f  oracle %.6f' % (k, out[k], ref[k]))
        # a flipped symbol moves the rate terms by ~1 bit / (N C h w) = 4e-5: allow 2e-4 then
        tol = 1e-4 if not mism.any() else 3e-4
        assert abs(out[k] - ref[k]) <= tol * max(1.0, abs(ref[k])), k
    G = tr.gradients()
    errs = []
    for name, g_ref in ref['grads'].items():
        g = G[name].astype(np.float64)
        w = np.asarray(Wt[name], np.float64)
        if name.startswith('autoencoder/') and name.endswith('/weights'):
            g = g + ae_cfg.regularization_factor * w
        elif name.endswith('/centers'):
            g = g + ae_cfg.regularization_factor_centers * w
        errs.append((float(np.linalg.norm(g - g_ref) / max(np.linalg.norm(g_ref), 1e-30)), name))
    errs.sort(reverse=True)
    for e, name in errs[:5]:
        print('  grad err %.2e  %s' % (e, name))
    med = errs[len(errs) // 2][0]
    q25 = errs[(3 * len(errs)) // 4][0]              # errs is sorted descending: the quartile of the BEST variables
    print('  gradient error over %d variables: median %.2e, best quartile %.2e' % (len(errs), med, q25))
    # float32-class kernels against float64 through ~70 conv+BN layers.  A pre-activation that lands on the other side of
    # zero than in float64 is a discrete event (tests/test_gpu_training_step.py): it moves the gradient of its own layer by
    # ~1e-2 and of EVERYTHING upstream of it (here: more than half of the 219 variables when it happens early in the
    # decoder) by ~5e-4, whichever float32 implementation runs -- so the median is gated at 1e-3, and the precision of
    # the kernels themselves is read off the variables downstream of any such event: best quartile 1.4e-6 in float32
    # mode (gate 5e-5).  The tensor-core mode measures 2.1e-4 there (gate 5e-4): its gradient tensors go through ONE
    # power-of-two scale per tensor before the fp16 hi/lo split, so elements far below the tensor's maximum keep fewer
    # than 22 bits -- float32-class for the forward pass, ~12 bits worse for gradients.
    assert med < 1e-3
    assert q25 < (5e-5 if mode == 'fp32' else 5e-4)
    assert errs[0][0] < 2e-2
