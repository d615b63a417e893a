"""tcgen05 3x3 conv (csrc/conv_tc.cu) against the float32 FFMA kernel, layer by layer,
through the C-ABI test hook ic_debug_conv3x3; then the whole encoder in EXACT / FAST mode."""
import numpy as np
import pytest
import torch

from conftest import load_golden, symbol_margin
from oracle import imgcomp_oracle as O

pytestmark = pytest.mark.gpu


def _conv(ae, decoder, layer, x, r1, r2, mode):
    from imgcomp_cvpr_b200 import _lib
    L = _lib.lib()
    N, H, W, _ = x.shape
    out = torch.full_like(x, float('nan'))
    ws = torch.empty(4 * x.numel() * 4 + 4096, dtype=torch.uint8, device='cuda')
    _lib.check(L.ic_debug_conv3x3(ae._handle, decoder, layer, _lib.ptr(x), _lib.ptr(r1), _lib.ptr(r2), N, H, W,
                                  _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.MODES[mode], _lib.stream_ptr()))
    return out


@pytest.mark.parametrize('shape', [(1, 16, 8), (1, 16, 16), (2, 40, 40), (3, 48, 24), (1, 8, 8), (1, 17, 9), (2, 192, 128)])
@pytest.mark.parametrize('mode', ['exact', 'fast'])
def test_layer_matches_ffma(shape, mode, gpu_models):
    ae, pc, W = gpu_models('cvpr/low')
    N, H, Wd = shape
    g = torch.Generator(device='cuda').manual_seed(N * 1000 + H * 10 + Wd)
    x, r1, r2 = (torch.randn((N, H, Wd, 128), device='cuda', generator=g) * s for s in (1.0, 3.0, 0.5))
    for decoder, layer, use_res in ((0, 0, False), (0, 5, True), (1, 31, True)):
        a, b = (r1, r2) if use_res else (None, None)
        ref = _conv(ae, decoder, layer, x, a, b, 'fp32')
        out = _conv(ae, decoder, layer, x, a, b, mode)
        assert not torch.isnan(out).any()
        scale = max(1.0, ref.abs().max().item())
        err = (out - ref).abs().max().item()
        # exact: float32-class (fp16x3 split, fp32 accumulate in TMEM); fast: one fp16 pass
        assert err < (1e-4 if mode == 'exact' else 2e-2) * scale, (decoder, layer, err, scale)


def test_layer_against_float64(gpu_models):
    """EXACT mode vs a float64 convolution: same error class as the FFMA float32 kernel."""
    ae, pc, W = gpu_models('cvpr/low')
    g = torch.Generator(device='cuda').manual_seed(7)
    x = torch.randn((1, 32, 32, 128), device='cuda', generator=g)
    scope = 'autoencoder/encoder/res_block_enc_0/enc_0_1/conv1'
    w = torch.from_numpy(W[scope + '/weights']).double().cuda().permute(3, 2, 0, 1)
    # float64 reference on the CPU: the first float64 cuDNN convolution of a process can spend minutes in cuDNN's own set-up
    y = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2).cpu(), w.cpu(), padding=1).cuda()
    bn = {k: torch.from_numpy(W[scope + '/BatchNorm/' + k]).double().cuda() for k in
          ('gamma', 'beta', 'moving_mean', 'moving_variance')}
    y = (y - bn['moving_mean'][None, :, None, None]) / torch.sqrt(bn['moving_variance'] + 1e-5)[None, :, None, None] \
        * bn['gamma'][None, :, None, None] + bn['beta'][None, :, None, None]
    y = torch.relu(y).permute(0, 2, 3, 1)
    e_ffma = (_conv(ae, 0, 0, x, None, None, 'fp32').double() - y).abs().mean().item()
    e_tc = (_conv(ae, 0, 0, x, None, None, 'exact').double() - y).abs().mean().item()
    e_fast = (_conv(ae, 0, 0, x, None, None, 'fast').double() - y).abs().mean().item()
    print('mean |err| vs float64: ffma %.3e  tc-exact %.3e  tc-fast %.3e' % (e_ffma, e_tc, e_fast))
    assert e_tc < 1e-5 and e_tc < 20 * max(e_ffma, 1e-8)
    assert e_fast < 1e-2


def test_accumulate_gain_against_float64(synth, monkeypatch):
    """The epilogue multiplies the main (hi*hi) accumulator by 1 + 1.606e-8 * accumulate steps to undo the MEAN of the
    tensor core's round-toward-zero fp32 accumulate (conv_tc.cu: launch_cat; IC_TC_ACC_GAIN=0 switches it off).  The
    correction is exact in the mean when the running sum keeps the sign of the final sum and over-corrects by at most its
    own size (1.2e-6 relative for a 3x3 layer) when it does not.  Measured against float64 on a layer without ReLU (the
    accumulator can be backed out of the output) for sign-mixed inputs (randn x 20), one-sided inputs (ReLU'd) and a
    large-offset input: the signed mean error must shrink and the mean magnitude may not grow by more than 1e-6."""
    from imgcomp_cvpr_b200 import autoencoder
    a, p, W = synth('cvpr/low')
    scope = 'autoencoder/encoder/res_block_enc_0/enc_0_1/conv2'         # layer index 1: activation_fn=None
    w = torch.from_numpy(W[scope + '/weights']).double().cuda().permute(3, 2, 0, 1)
    bn = {k: torch.from_numpy(W[scope + '/BatchNorm/' + k]).double().cuda() for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')}
    sc = bn['gamma'] / torch.sqrt(bn['moving_variance'] + 1e-5)
    g = torch.Generator(device='cuda').manual_seed(11)
    inputs = {'mixed': torch.randn((1, 64, 64, 128), device='cuda', generator=g) * 20,
              'relu': torch.relu(torch.randn((1, 64, 64, 128), device='cuda', generator=g)) * 1.5,
              'offset': torch.randn((1, 64, 64, 128), device='cuda', generator=g) * 2 + 30}
    res = {}
    for gain in ('1', '0'):
        monkeypatch.setenv('IC_TC_ACC_GAIN', gain)
        ae = autoencoder.get_network_cls(a)(a, weights=W, mode='exact')
        for name, x in inputs.items():
            pre = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2).cpu(), w.cpu(), padding=1).cuda().permute(0, 2, 3, 1)
            out = _conv(ae, 0, 1, x, None, None, 'exact').double()
            acc = (out - bn['beta']) / sc + bn['moving_mean']
            m = pre.abs() > 0.2 * pre.pow(2).mean().sqrt()
            rel = ((acc - pre) / pre)[m]
            res[(name, gain)] = (rel.mean().item(), rel.abs().mean().item())
        del ae
    for name in inputs:
        on, off = res[(name, '1')], res[(name, '0')]
        print('%-6s exact vs float64, (mean signed rel err, mean |rel err|): compensation on (%.2e, %.2e)  off (%.2e, %.2e)' % (
            (name,) + on + off))
        assert abs(on[0]) <= abs(off[0]) + 2e-7, name
        assert on[1] <= off[1] + 1e-6, name


@pytest.mark.parametrize('name,ae_name', [('cfg1_low_1x128x128', 'cvpr/low'), ('ragged_low_2x48x72', 'cvpr/low'),
                                          ('tiny_hi_1x40x24', 'cvpr/hi')])
def test_fast_mode_flip_rate_is_reported(name, ae_name, gpu_models):
    """FAST mode is NOT the parity mode: it is reported with its symbol flip rate."""
    ae, pc, W = gpu_models(ae_name, 'fast')
    g = load_golden(name)
    enc = ae.encode(torch.from_numpy(g['x_u8']).cuda(), False)
    flips = (enc.symbols.cpu().numpy() != g['symbols']).mean()
    print('fast mode: %s symbol flip rate %.4f, max |dz| %.3e' % (name, flips, np.abs(enc.z.cpu().numpy() - g['z']).max()))
    assert flips < 0.05


@pytest.mark.parametrize('name,ae_name', [('cfg1_low_1x128x128', 'cvpr/low'), ('ragged_low_2x48x72', 'cvpr/low'),
                                          ('tiny_hi_1x40x24', 'cvpr/hi')])
def test_decoder_tensor_core_transposed_convs(name, ae_name, gpu_models):
    """decode in EXACT mode (from_bn / h12 / h13 as depth-to-space tcgen05 convs) vs the FFMA decoder
    on the same qhard, and vs the reference-run golden image."""
    ae32, _, W = gpu_models(ae_name, 'fp32')
    aex, _, _ = gpu_models(ae_name, 'exact')
    g = load_golden(name)
    q = torch.from_numpy(W['autoencoder/encoder/centers'][g['symbols'].astype(np.int64)]).cuda()
    x32 = ae32.decode(q, False).clone()
    u32 = ae32.extra['x_out_u8'].clone()
    xex = aex.decode(q, False)
    uex = aex.extra['x_out_u8']
    d = (xex - x32).abs().max().item()
    print('%s: decode exact vs fp32 max |dx| %.3e, uint8 mismatches %.2e' % (name, d, (uex != u32).float().mean().item()))
    assert d < 2e-2
    assert (uex != u32).float().mean().item() < 2e-3
    assert (uex.cpu().numpy() != g['x_out_u8']).mean() < 2e-3
    assert torch.equal(uex, xex.to(torch.uint8))


@pytest.mark.parametrize('shape', [(1, 8, 8), (2, 48, 72), (1, 128, 128), (3, 40, 24), (1, 256, 512)])
@pytest.mark.parametrize('mode', ['exact', 'fast'])
def test_dedicated_h1_kernel_against_generic_path_and_oracle(shape, mode, synth, monkeypatch):
    """conv_h1.cu (normalise + im2col + 5x5 stride-2 conv from the uint8 / float32 image in one kernel) against the
    round-1 path (prep pass + grouped-tap kernel, IC_H1_GENERIC=1) through the whole encoder, on shapes with partial
    16 x 8 tiles, and for uint8 and float32 input."""
    from imgcomp_cvpr_b200 import autoencoder, weights as wm
    a, p, W = synth('cvpr/low')
    N, H, Wd = shape
    x = torch.from_numpy(wm.synthetic_images(N, H, Wd, seed=H + Wd)).cuda()
    ae = autoencoder.get_network_cls(a)(a, weights=W, mode=mode)
    monkeypatch.setenv('IC_H1_GENERIC', '1')
    z_old = ae.encode(x, False).z.clone()
    monkeypatch.setenv('IC_H1_GENERIC', '0')
    e = ae.encode(x, False)
    z_new, s_new = e.z.clone(), e.symbols.clone()
    z_f = ae.encode(x.float(), False).z
    assert torch.equal(z_f, z_new)                                   # tf.to_float fused identically
    tol = 3e-4 if mode == 'exact' else 0.15          # fast = single fp16 pass: two roundings of the same sum differ by ~1e-2
    assert float((z_new - z_old).abs().max()) < tol
    if mode == 'exact':
        O.set_backend('torch')
        try:
            ref = O.encode(x.cpu().numpy().astype(np.float32), W, 32)
            z64 = O.encode(x.cpu().numpy().astype(np.float64), W, 32, dtype=np.float64)['z']
        finally:
            O.set_backend('numpy')
        np.testing.assert_allclose(z_new.cpu().numpy(), ref['z'], atol=1e-3)
        safe = symbol_margin(z64, W['autoencoder/encoder/centers']) > 6e-4
        assert (s_new.cpu().numpy()[safe] == ref['symbols'][safe]).all()


@pytest.mark.parametrize('cat', ['1', '2'])
def test_b_concatenated_3x3_kernel_opt_in(cat, synth, monkeypatch):
    """IC_CONV_CAT=1 (one CTA) / =2 (CTA pairs, cta_group::2): the B-concatenated 3x3 kernel with separate hi*hi / cross-term
    accumulators and skewed tiles (conv_tc.cu: conv_cat_kernel).  Opt-in because it measures slower than the default
    kernel; it must still be right: whole encoder against the default kernel and the float64 oracle."""
    from imgcomp_cvpr_b200 import autoencoder, weights as wm
    a, p, W = synth('cvpr/low')
    x = torch.from_numpy(wm.synthetic_images(2, 192, 160, seed=9)).cuda()
    monkeypatch.setenv('IC_CONV_CAT', '0')
    ae0 = autoencoder.get_network_cls(a)(a, weights=W, mode='exact')
    monkeypatch.setenv('IC_CONV_CAT', cat)
    ae1 = autoencoder.get_network_cls(a)(a, weights=W, mode='exact')
    z0 = ae0.encode(x, False).z.clone()
    e1 = ae1.encode(x, False)
    z1, s1 = e1.z.clone(), e1.symbols.clone()
    assert torch.equal(ae1.encode(x, False).z, z1)                      # deterministic
    assert float((z1 - z0).abs().max()) < 5e-4
    O.set_backend('torch')
    try:
        ref = O.encode(x.cpu().numpy().astype(np.float32), W, 32)
        z64 = O.encode(x.cpu().numpy().astype(np.float64), W, 32, dtype=np.float64)['z']
    finally:
        O.set_backend('numpy')
    d0, d1 = np.abs(z0.cpu().numpy() - z64), np.abs(z1.cpu().numpy() - z64)
    print('IC_CONV_CAT=%s: mean |z - z64| %.2e (default kernel %.2e), max %.2e (%.2e)' % (cat, d1.mean(), d0.mean(), d1.max(), d0.max()))
    safe = symbol_margin(z64, W['autoencoder/encoder/centers']) > 2e-4
    assert (s1.cpu().numpy()[safe] == ref['symbols'][safe]).all()
    assert d1.mean() <= 1.2 * d0.mean()                                  # separate accumulators: not less accurate
