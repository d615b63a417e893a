"""GPU parity tests of the hot path (run with -m gpu on the B200 box).

Every check calls the product through its public host API, i.e. through the
C ABI of libimgcomp_b200.so, and compares with
  * the golden vectors produced by the reference's own modules (tests/golden), and
  * the numpy oracle (oracle/imgcomp_oracle.py) on the same seeded inputs.
Tolerances: symbols bit-exact wherever the float64 latent is further than
EPS_MARGIN from a quantizer decision boundary (two float32 implementations of a
34-layer conv stack differ by summation order), bpp and MS-SSIM
within 1e-4 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden, symbol_margin
from oracle import imgcomp_oracle as O

pytestmark = pytest.mark.gpu
MODES = ['fp32', 'exact']
# (z atol, symbol margin eps): 'fp32' = float32 FFMA kernels, 'exact' = tcgen05 fp16x3 (fp32 accumulate in the
# tensor core truncates, so its error is a few times that of an FFMA chain -- DESIGN.md "Arithmetic")
TOL = {'fp32': (3e-4, 1e-4), 'exact': (1e-3, 6e-4)}


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', sorted(GOLDEN_CASES))
def test_val_graph_against_reference_goldens(name, mode, gpu_models):
    """encode -> decode(qhard) -> bitcost(qbar) -> bpp / MS-SSIM, as code/val.py:81-94."""
    from imgcomp_cvpr_b200 import bits, ms_ssim, ms_ssim_np
    ae_name, _ = GOLDEN_CASES[name]
    ae, pc, W = gpu_models(ae_name, mode)
    g = load_golden(name)
    x_u8 = _cuda(g['x_u8'])
    enc = ae.encode(x_u8, is_training=False)
    x_out = ae.decode(enc.qhard, is_training=False)
    bc = pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
    sym = enc.symbols.cpu().numpy()
    z = enc.z.cpu().numpy()
    z_atol, eps_margin = TOL[mode]
    z64 = O.encode(g['x_u8'].astype(np.float64), W, ae.config.num_chan_bn, dtype=np.float64)['z']
    mism = sym != g['symbols']
    print('%s/%s: max|z-golden| %.2e  max|z-z64| %.2e  symbol mismatches %d/%d' % (
        name, mode, np.abs(z - g['z']).max(), np.abs(z - z64).max(), mism.sum(), mism.size))
    np.testing.assert_allclose(z, g['z'], atol=z_atol, rtol=0)
    safe = symbol_margin(z64, W['autoencoder/encoder/centers']) > eps_margin
    assert (sym[safe] == g['symbols'][safe]).all(), 'symbol flipped away from a decision boundary'
    assert mism.mean() <= (2e-4 if mode == 'fp32' else 1e-3), mism.sum()
    same = ~mism
    assert np.array_equal(enc.symbols.cpu().numpy(), ae.extra['symbols_u8'].cpu().numpy().astype(np.int64))
    centers = W['autoencoder/encoder/centers']
    assert np.array_equal(enc.qhard.cpu().numpy(), centers[sym])          # qhard = centers[symbols] exactly
    np.testing.assert_allclose(enc.qbar.cpu().numpy()[same], g['qbar'][same], atol=1e-6)
    if 'heatmap' in g:
        np.testing.assert_allclose(enc.heatmap.cpu().numpy(), g['heatmap'], atol=z_atol)
    np.testing.assert_allclose(bc.cpu().numpy()[same], g['bitcost'][same], atol=3e-3)
    # per image bpp (val.py runs batch = 1)
    for i in range(x_u8.shape[0]):
        bpp = bits.bitcost_to_bpp(bc[i:i + 1], x_u8[i:i + 1]).item()
        print('   bpp %.6f golden %.6f' % (bpp, g['bpp'][i]))
        assert abs(bpp - g['bpp'][i]) < 1e-4
        assert abs(pc.last_bits_per_image[i].item() / (x_u8.shape[2] * x_u8.shape[3]) - g['bpp'][i]) < 1e-4
    if 'x_out' in g:
        np.testing.assert_allclose(x_out.cpu().numpy(), g['x_out'], atol=1e-2)
    x_out_u8 = ae.extra['x_out_u8']
    assert (x_out_u8.cpu().numpy() != g['x_out_u8']).mean() < 2e-3
    assert torch.equal(x_out_u8, x_out.to(torch.uint8))                   # tf.cast truncation (val.py:91)
    ms = ms_ssim_np.MultiScaleSSIM_batch(x_u8, x_out_u8, data_format='NCHW').cpu().numpy()
    np.testing.assert_allclose(ms, g['ms_ssim_np'], atol=1e-4)
    if 'ms_ssim_tf_raises' in g:
        with pytest.raises(RuntimeError):
            ms_ssim.MultiScaleSSIM(x_u8.float(), x_out, data_format='NCHW')
    else:
        v = ms_ssim.MultiScaleSSIM(x_u8.float(), x_out, data_format='NCHW').item()
        assert abs(v - float(g['ms_ssim_tf'])) < 1e-4


def test_quantizer_bit_exact_on_identical_input(gpu_models):
    """quantizer.quantize on the same z: symbols and qhard bit-exact incl. exact
    midpoints (ties -> first index) and exact centres; qsoft to float32 rounding."""
    from imgcomp_cvpr_b200 import quantizer
    ae, pc, W = gpu_models('cvpr/low')
    centers = W['autoencoder/encoder/centers']
    g = load_golden('cfg1_low_1x128x128')
    rng = np.random.RandomState(3)
    cs = np.sort(centers)
    mids = ((cs[1:] + cs[:-1]) / 2).astype(np.float32)
    special = np.concatenate([centers, mids, np.nextafter(mids, np.float32(10)), np.nextafter(mids, np.float32(-10)),
                              np.float32([0, -0.0, 5, -5, 1e-30, 100, -100])])
    z = np.concatenate([g['z'].ravel(), rng.uniform(-2.5, 2.5, 100000).astype(np.float32), special.astype(np.float32)])
    z = np.resize(z, (1, 4, 128, z.size // 512 + 1)).astype(np.float32)
    qsoft, qhard, sym = quantizer.quantize(_cuda(z), _cuda(centers), sigma=1)
    osoft, ohard, osym = O.quantize(z, centers, 1)
    assert np.array_equal(sym.cpu().numpy(), osym)
    assert np.array_equal(qhard.cpu().numpy(), ohard)
    np.testing.assert_allclose(qsoft.cpu().numpy(), osoft, atol=5e-7)
    with pytest.raises(AssertionError):
        quantizer.quantize(_cuda(z).double(), _cuda(centers), 1)
    with pytest.raises(AssertionError):
        quantizer.quantize(_cuda(z)[0], _cuda(centers), 1)


def test_probclass_batched_freqs_match_reference_loop_and_are_causal(gpu_models):
    ae, pc, W = gpu_models('cvpr/low')
    from imgcomp_cvpr_b200 import probclass
    g = load_golden('tiny_low_1x64x64')
    syms = g['symbols'][0].astype(np.int64)
    centers = _cuda(W['autoencoder/encoder/centers'])
    f, bits = pc.freqs(_cuda(syms)[None], centers)
    f = f[0].cpu().numpy()
    assert f.shape == g['freqs'].shape and f.min() >= 1
    # float32 softmax * 1e9 truncated: hi/lo tensor-core logits differ from the float32 chain by ~1e-5
    assert np.abs(f - g['freqs']).max() <= 3e4
    assert abs(bits.item() - float(g['theory_bits'])) < 0.5
    # the per-context network of the reference (PredictionNetwork.get_freqs) is bit-identical
    # to the batched pass at the same position
    ae.encode(_cuda(g['x_u8']), False)
    pred = probclass.PredictionNetwork(pc, pc.config, ae.get_centers_variable(), None)
    sp = pred.pad_symbols_volume(syms)
    assert pred.input_ctx_shape == (5, 9, 9)
    for (c, y, x) in ((0, 0, 0), (0, 0, 1), (5, 3, 2), (31, 7, 7), (12, 0, 7), (31, 0, 0)):
        fc = pred.get_freqs(sp[c:c + 5, y:y + 9, x:x + 9])
        assert np.array_equal(fc, f[c, y, x]), (c, y, x)
        pr = pred.get_pr(sp[c:c + 5, y:y + 9, x:x + 9])
        np.testing.assert_allclose(pr, f[c, y, x] / 1e9, atol=2e-6)
    # pc.logits on the manually padded volume (code/probclass.py:130-135) == per-context logits, bit for bit
    qpad = torch.from_numpy(W['autoencoder/encoder/centers'][sp]).cuda()[None]
    full = pc.logits(qpad)[0]
    assert tuple(full.shape) == syms.shape + (6,)
    for (c, y, x) in ((0, 0, 0), (5, 3, 2), (31, 7, 7)):
        one = pc.logits(qpad[:, c:c + 5, y:y + 9, x:x + 9].contiguous())[0, 0, 0, 0]
        assert torch.equal(one, full[c, y, x])
    # causality: tables before position p do not change when symbols at/after p change
    rng = np.random.RandomState(0)
    flat = syms.reshape(-1).copy()
    p = flat.size // 2 + 13
    flat[p:] = rng.randint(0, 6, flat.size - p)
    f2, _ = pc.freqs(_cuda(flat.reshape(syms.shape))[None], centers)
    f2 = f2[0].cpu().numpy().reshape(-1, 6)
    assert np.array_equal(f2[:p + 1], f.reshape(-1, 6)[:p + 1])
    assert not np.array_equal(f2[p + 1:], f.reshape(-1, 6)[p + 1:])


@pytest.mark.parametrize('shape', [(1, 4, 9, 9), (1, 32, 8, 8), (2, 9, 21, 37), (5, 32, 40, 24)])
def test_context_model_depth_walk_against_two_group_schedule(shape, gpu_models, monkeypatch):
    """csrc/conv_tc.cu, ConvTcParams::walk (the default of the inference layers): every input slice is fetched once and
    feeds the output slices on both sides of it, accumulators in a ring of four [X | Y] tiles.  Against the two-group
    schedule (IC_PC_WALK=0: each output slice loads its two input slices) on the same weights and symbols: same sums in
    another order, so logits / bit costs agree to float32 rounding, tables to a few hundred counts of 1e9; every table is
    positive and the walk is deterministic.  Shapes: fewer columns than CTAs (depth segments of 1..3 outputs), a ring
    that wraps several times, ragged tiles, and a batch with more columns than twice the CTA count (whole columns)."""
    ae, pc, W = gpu_models('cvpr/low')
    N, C, H, Wd = shape
    rng = np.random.RandomState(7)
    sym = _cuda(rng.randint(0, 6, size=(N, C, H, Wd)).astype(np.int64))
    centers = _cuda(W['autoencoder/encoder/centers'])
    q = centers[sym]
    f1, b1 = pc.freqs(sym, centers)
    bc1 = pc.bitcost(q, sym, False, pad_value=float(centers[0]))
    f1b, _ = pc.freqs(sym, centers)
    assert torch.equal(f1, f1b)
    monkeypatch.setenv('IC_PC_WALK', '0')
    f0, b0 = pc.freqs(sym, centers)
    bc0 = pc.bitcost(q, sym, False, pad_value=float(centers[0]))
    monkeypatch.delenv('IC_PC_WALK')
    assert int(f1.min()) >= 1
    df = int((f1 - f0).abs().max())
    db = float((bc1 - bc0).abs().max())
    print('  depth walk %s: max table difference %d / 1e9, max bit-cost difference %.2e' % (shape, df, db))
    assert df <= 2000, df
    assert db <= 2e-5 * max(1.0, float(bc0.abs().max())), db
    assert torch.allclose(b1, b0, rtol=1e-6, atol=1e-3)


def test_real_bpp_round_trip(gpu_models):
    """--real_bpp (val.py:161-175): coded bits ~ theoretical bits ~ loss bpp; stream decodes."""
    from imgcomp_cvpr_b200 import bit_counter, bpp_helpers, probclass
    ae, pc, W = gpu_models('cvpr/low')
    g = load_golden('tiny_low_1x64x64')
    x = _cuda(g['x_u8'])
    enc = ae.encode(x, False)
    bc = pc.bitcost(enc.qbar, enc.symbols, False, pad_value=pc.auto_pad_value(ae))
    pred = probclass.PredictionNetwork(pc, pc.config, ae.get_centers_variable(), None)
    checker = probclass.ProbclassNetworkTesting(pc, ae, None)
    fetcher = bpp_helpers.BppFetcher(pred, checker)
    sym = enc.symbols.cpu().numpy()
    num_pixels = bpp_helpers.num_pixels_in_image(g['x_u8'][0])
    bpp_real, bpp_theory = fetcher.get_bpp(sym, num_pixels)
    bpp_loss = bc.sum().item() / num_pixels
    assert abs(bpp_theory - bpp_loss) < 1e-3                      # val.py:174
    assert abs(bpp_real - bpp_theory) * num_pixels < 50           # bit_counter.py:51
    if np.array_equal(sym[0], g['symbols'][0]):
        assert abs(bpp_real * num_pixels - int(g['real_bits'])) <= 16


def test_msssim_pairs(gpu_models):
    from imgcomp_cvpr_b200 import ms_ssim, ms_ssim_np
    g = load_golden('msssim_pairs')
    for tag in 'abc':
        x, y = _cuda(g['x_' + tag]), _cuda(g['y_' + tag])
        v = ms_ssim_np.MultiScaleSSIM_batch(x, y).cpu().numpy()
        np.testing.assert_allclose(v, g['np_' + tag], atol=1e-9)
        t = ms_ssim.MultiScaleSSIM(x.float(), y.float(), data_format='NCHW').item()
        assert abs(t - float(g['tf_' + tag])) < 2e-5
        t2 = ms_ssim.MultiScaleSSIM(x.float().permute(0, 2, 3, 1), y.float().permute(0, 2, 3, 1)).item()
        assert t2 == t
    assert ms_ssim_np.tf_msssim_np(x, y, 'NCHW').dtype == torch.float32
    with pytest.raises(RuntimeError):
        ms_ssim.MultiScaleSSIM(x.float(), y.float()[:, :, :-8], data_format='NCHW')
    with pytest.raises(RuntimeError):
        ms_ssim.MultiScaleSSIM(x.float()[0], y.float()[0], data_format='NCHW')
    # identical images -> exactly 1
    assert abs(ms_ssim_np.MultiScaleSSIM_batch(x, x).item() - 1.0) < 1e-12


def test_msssim_stream_kernel_matches_tiled_kernel(gpu_models, monkeypatch):
    """row-streaming level kernel (11-tap levels without REFLECT padding) against the 16x16-tile kernel it replaces there:
    identical per-pixel arithmetic, only the (double) partial-sum order differs.  Odd sizes: ragged last blocks in x and y."""
    from imgcomp_cvpr_b200 import ms_ssim, ms_ssim_np
    rng = np.random.RandomState(7)
    for shape in ((4, 3, 200, 333), (1, 3, 176, 176), (3, 3, 305, 190), (2, 3, 400, 520)):
        x = rng.randint(0, 256, shape).astype(np.uint8)
        y = np.clip(x.astype(np.int32) + rng.randint(-12, 13, shape), 0, 255).astype(np.uint8)
        x, y = _cuda(x), _cuda(y)
        monkeypatch.delenv('IC_MSSSIM_TILED', raising=False)
        v_np = ms_ssim_np.MultiScaleSSIM_batch(x, y).cpu().numpy()
        v_tf = ms_ssim.MultiScaleSSIM(x.float(), y.float(), data_format='NCHW').item()
        monkeypatch.setenv('IC_MSSSIM_TILED', '1')
        w_np = ms_ssim_np.MultiScaleSSIM_batch(x, y).cpu().numpy()
        w_tf = ms_ssim.MultiScaleSSIM(x.float(), y.float(), data_format='NCHW').item()
        monkeypatch.delenv('IC_MSSSIM_TILED', raising=False)
        np.testing.assert_allclose(v_np, w_np, rtol=0, atol=1e-13)
        assert abs(v_tf - w_tf) < 1e-6


def test_res_shallow_64_context_model_against_oracle(gpu_models):
    """pc_configs/cvpr/res_shallow_64 (arch_param__k = 64, code/pc_configs/cvpr/res_shallow_64:8): bit cost, logits and coder
    tables of the 64-channel context model against the oracle (float32 chain on the FFMA kernels: no tensor-core packing
    exists for k = 64)."""
    ae, pc, W = gpu_models('cvpr/low', 'fp32', 'cvpr/res_shallow_64')
    assert W['probclass3d/logits/res1/conv3d_conv1_mask/weights'].shape == (2, 3, 3, 64, 64)
    rng = np.random.RandomState(11)
    sym = rng.randint(0, 6, (2, 32, 7, 9)).astype(np.int64)
    centers = W['autoencoder/encoder/centers'].astype(np.float32)
    q = centers[sym]
    bc = pc.bitcost(_cuda(q), _cuda(sym), False, pad_value=float(centers[0])).cpu().numpy()
    bc_ref, logits_ref = O.pc_bitcost(q, sym, W, centers[0], np.float64)
    np.testing.assert_allclose(bc, bc_ref, atol=2e-4, rtol=1e-4)
    assert abs(bc.sum() / bc_ref.sum() - 1) < 1e-5                                 # bpp: 1e-4 is the north-star bound
    f, bits = pc.freqs(_cuda(sym[:1]), _cuda(centers), codec=True)
    f_ref = O.pc_freqs_volume(sym[0], W, centers)
    assert np.abs(f[0].cpu().numpy() - f_ref).max() <= 256                          # float32 softmax * 1e9, like the k = 24 tables
    qpad = _cuda(O.pad_for_probclass3d(q, 9, centers[0]).astype(np.float32))
    lg = pc.logits(qpad).cpu().numpy()
    np.testing.assert_allclose(lg, logits_ref, atol=5e-4)


def test_api_errors_mirror_reference(gpu_models, synth):
    from imgcomp_cvpr_b200 import autoencoder, probclass
    a, p, W = synth('cvpr/low')
    ae = autoencoder.get_network_cls(a)(a, weights=W)
    with pytest.raises(ValueError):
        ae.get_centers_variable()                                  # autoencoder.py:66-67
    x = torch.zeros((1, 3, 36, 64), dtype=torch.uint8, device='cuda')
    with pytest.raises(ValueError):
        ae.encode(x, False)                                        # not a multiple of 8
    with pytest.raises(AssertionError):
        ae.encode(x.double(), False)                               # autoencoder.py:51
    with pytest.raises(RuntimeError):
        autoencoder.get_network_cls(a)(a).encode(x[:, :, :32], False)     # no weights
    pc = probclass.get_network_cls(p)(p, num_centers=6, weights=W)
    with pytest.raises(AssertionError):
        pc.logits(torch.zeros((1, 5, 9, 9, 1), device='cuda'))     # bitcost first (probclass.py:132)
    assert ae.get_subsampling_factor() == 8
    assert pc.get_context_size(p) == 9 and pc.get_context_shape(p) == (5, 9, 9)
    with pytest.raises(KeyError):
        autoencoder.get_network_cls(type('C', (), {'arch': 'TwoLayerNet'}))


def test_batch_and_position_independence(gpu_models):
    """Size-independent properties at a larger shape: an image encodes to the same bits
    alone or inside a batch, runs are deterministic, and float32 vs uint8 input agree."""
    from imgcomp_cvpr_b200 import weights as wm
    ae, pc, W = gpu_models('cvpr/low')
    x = _cuda(wm.synthetic_images(3, 96, 160, seed=5))
    e_all = ae.encode(x, False)
    sym_all, z_all = e_all.symbols.clone(), e_all.z.clone()
    e1 = ae.encode(x[1:2].contiguous(), False)
    assert torch.equal(e1.symbols, sym_all[1:2]) and torch.equal(e1.z, z_all[1:2])
    e_f = ae.encode(x.float(), False)
    assert torch.equal(e_f.symbols, sym_all) and torch.equal(e_f.z, z_all)
    bc_all = pc.bitcost(e_f.qbar, e_f.symbols, False, pad_value=pc.auto_pad_value(ae)).clone()
    bc1 = pc.bitcost(e1.qbar, e1.symbols, False, pad_value=pc.auto_pad_value(ae))
    assert torch.equal(bc1, bc_all[1:2])
    xo = ae.decode(e_f.qhard, False).clone()
    assert torch.equal(ae.decode(e1.qhard, False), xo[1:2])
    assert xo.min().item() >= 0 and xo.max().item() <= 255


def test_val_driver_matches_goldens(gpu_models):
    """imgcomp_cvpr_b200.val.validate (padding, same-size batching, metric aggregation) on the golden images."""
    from imgcomp_cvpr_b200 import val
    ae, pc, W = gpu_models('cvpr/low')
    g1, g2 = load_golden('tiny_low_1x64x64'), load_golden('ragged_low_2x48x72')
    imgs = [g1['x_u8'][0], g2['x_u8'][0], np.transpose(g2['x_u8'][1], (1, 2, 0))]    # CHW and HWC accepted
    avgs, n, rows = val.validate(imgs, ae, pc, real_bpp=False, batch_size=2)
    assert n == 3
    want_bpp = [g1['bpp'][0], g2['bpp'][0], g2['bpp'][1]]
    want_ms = [g1['ms_ssim_np'][0], g2['ms_ssim_np'][0], g2['ms_ssim_np'][1]]
    for r, b, m in zip(rows, want_bpp, want_ms):
        assert abs(r['bpp'] - b) < 1e-4 and abs(r['ms-ssim'] - m) < 1e-4
    assert abs(avgs['bpp'] - np.mean(want_bpp)) < 1e-4
    # an image that needs padding: 60x70 -> 64x72, centred (images_iterator.py:39-59)
    odd = np.transpose(g2['x_u8'][0], (1, 2, 0))[:44, :70]
    _, _, rows2 = val.validate([odd], ae, pc)
    assert np.isfinite(rows2[0]['bpp'])
    rows3 = val.measure_batch(torch.from_numpy(g1['x_u8']).cuda(), ae, pc, real_bpp=True)
    assert abs(rows3[0]['bpp_real'] - rows3[0]['bpp_theory']) * 64 * 64 < 50


@pytest.mark.parametrize('shape', [(1, 8, 8), (2, 16, 8), (1, 8, 40), (3, 24, 136)])
def test_smallest_and_odd_tile_shapes_all_modes(shape, gpu_models):
    """Edge shapes (one latent pixel, partial tensor-core tiles, W/4 <= 8): exact mode vs the float32 path vs the oracle."""
    from imgcomp_cvpr_b200 import weights as wm
    N, H, Wd = shape
    x = wm.synthetic_images(N, H, Wd, seed=H * 100 + Wd)
    xc = _cuda(x)
    ae32, pc, W = gpu_models('cvpr/low', 'fp32')
    aex, _, _ = gpu_models('cvpr/low', 'exact')
    e32 = ae32.encode(xc, False)
    z32, s32 = e32.z.clone(), e32.symbols.clone()
    eex = aex.encode(xc, False)
    ref = O.encode(x.astype(np.float32), W, 32)
    np.testing.assert_allclose(z32.cpu().numpy(), ref['z'], atol=3e-4)
    np.testing.assert_allclose(eex.z.cpu().numpy(), ref['z'], atol=1e-3)
    z64 = O.encode(x.astype(np.float64), W, 32, dtype=np.float64)['z']
    safe = symbol_margin(z64, W['autoencoder/encoder/centers']) > 6e-4
    assert (s32.cpu().numpy()[safe] == ref['symbols'][safe]).all()
    assert (eex.symbols.cpu().numpy()[safe] == ref['symbols'][safe]).all()
    bc = pc.bitcost(e32.qbar, e32.symbols, False, pad_value=pc.auto_pad_value(ae32))
    obc, _ = O.pc_bitcost(ref['qbar'], ref['symbols'], W, W['autoencoder/encoder/centers'][0])
    same = s32.cpu().numpy() == ref['symbols']
    np.testing.assert_allclose(bc.cpu().numpy()[same], obc[same], atol=3e-3)
    xo32 = ae32.decode(e32.qhard, False).clone()
    xoex = aex.decode(e32.qhard, False)
    oref = O.decode(e32.qhard.cpu().numpy(), W)
    np.testing.assert_allclose(xo32.cpu().numpy(), oref, atol=2e-2)
    np.testing.assert_allclose(xoex.cpu().numpy(), oref, atol=5e-2)


def test_full_size_properties_kodak_shape(gpu_models):
    """BASELINE.json configs[1] shape (768x512, 2 images here): size-independent properties -- determinism,
    batch independence in exact mode, symbols in range, qhard = centres[symbols], bits finite and positive,
    coded size ~ theoretical size (the reference's own --real_bpp asserts), decode range."""
    from imgcomp_cvpr_b200 import bit_counter, probclass, weights as wm
    ae, pc, W = gpu_models('cvpr/low', 'exact')
    x = _cuda(wm.synthetic_images(2, 768, 512, seed=77))
    e = ae.encode(x, False)
    sym, z, qhard, qbar = e.symbols.clone(), e.z.clone(), e.qhard.clone(), e.qbar.clone()
    e2 = ae.encode(x, False)
    assert torch.equal(e2.symbols, sym) and torch.equal(e2.z, z)                    # deterministic
    e1 = ae.encode(x[1:2].contiguous(), False)
    assert torch.equal(e1.symbols, sym[1:2]) and torch.equal(e1.z, z[1:2])           # batch independent
    assert sym.min().item() >= 0 and sym.max().item() <= 5
    centers = torch.from_numpy(W['autoencoder/encoder/centers']).cuda()
    assert torch.equal(qhard, centers[sym])
    assert (qbar - qhard).abs().max().item() <= 4e-7                                 # qbar = qsoft + (qhard - qsoft)
    bc = pc.bitcost(qbar, sym, False, pad_value=pc.auto_pad_value(ae))
    assert torch.isfinite(bc).all() and bc.min().item() >= 0
    bits_img = pc.last_bits_per_image.clone()
    np.testing.assert_allclose(bits_img.cpu().numpy(), bc.double().sum(dim=(1, 2, 3)).cpu().numpy(), rtol=1e-6)
    pred = probclass.PredictionNetwork(pc, pc.config, ae.get_centers_variable(), None)
    nbits = bit_counter.encode_decode_to_file_ctx(sym[0].cpu().numpy(), pred, syms_format='CHW')
    f, theory = pred.get_all_freqs(sym[0])
    assert abs(nbits - theory) < 50 and abs(theory - bits_img[0].item()) < 5e-3 * theory + 1
    xo = ae.decode(qhard, False)
    assert xo.min().item() >= 0 and xo.max().item() <= 255 and torch.isfinite(xo).all()


def test_training_loss_forward(gpu_models, synth):
    """train.get_loss + Distortions, forward (code/train.py:303-336,352-431) against the oracle."""
    from imgcomp_cvpr_b200 import train
    a, p, W = synth('cvpr/low')
    ae, pc, _ = gpu_models('cvpr/low', 'fp32')
    g = load_golden('ragged_low_2x48x72')
    x = _cuda(g['x_u8']).float()
    enc = ae.encode(x, False)
    x_out = ae.decode(enc.qbar, False)
    bc = pc.bitcost(enc.qbar, enc.symbols, False, pad_value=pc.auto_pad_value(ae))
    reg = train.regularization_losses(a, p, W)
    assert reg[2] is None and reg[0] > 0 and reg[1] > 0
    a_psnr = type(a)(**dict(a.__dict__, distortion_to_minimize='psnr'))
    with pytest.raises(RuntimeError):                      # ms_ssim on this ragged shape raises, as in the reference
        train.Distortions(a, x, x_out, False)
    for is_training in (False,):
        d = train.Distortions(a_psnr, x, x_out, is_training)
        xo = x_out.cpu().numpy()
        xi = g['x_u8'].astype(np.float32)
        np.testing.assert_allclose(d.mse, O.mse_per_img(xi, xo, True).mean(), rtol=1e-5)
        np.testing.assert_allclose(d.psnr, O.psnr_per_img(xi, xo, True).mean(), rtol=1e-5)
        # ms_ssim on a tiny ragged shape raises in the reference; use mse-minimising config for it
    a_mse = type(a)(**dict(a.__dict__, distortion_to_minimize='mse'))
    d = train.Distortions(a_mse, x, x_out, True)          # training + mse: NO int cast
    np.testing.assert_allclose(d.mse, O.mse_per_img(g['x_u8'].astype(np.float32), x_out.cpu().numpy(), False).mean(), rtol=1e-5)
    total, H_real, pc_comps, ae_comps = train.get_loss(a, ae, pc, d.d_loss_scaled, bc, enc.heatmap, reg)
    o_total, o_real, o_mask, o_pcl = O.get_loss(bc.cpu().numpy(), enc.heatmap.cpu().numpy(), d.d_loss_scaled, a.H_target,
                                                a.beta, reg[0], reg[1], 0.0)
    assert abs(H_real - o_real) < 1e-5 and abs(dict(pc_comps)['H_mask'] - o_mask) < 1e-5
    assert abs(dict(pc_comps)['pc_loss'] - o_pcl) < 1e-2 and abs(total - o_total) < 1e-3 * abs(o_total)
    x160 = torch.rand((2, 3, 160, 160), device='cuda') * 255
    y160 = (x160 + torch.randn_like(x160) * 4).clamp(0, 255)
    dm = train.Distortions(a, x160, y160, True)           # ms_ssim config on the training crop size
    assert abs(dm.d_loss_scaled - a.K_ms_ssim * (1 - dm.ms_ssim)) < 1e-3
    assert abs(dm.ms_ssim - O.ms_ssim_tf(x160.cpu().numpy(), y160.cpu().numpy())[0]) < 1e-4


def test_training_mode_boundary_drops_in(gpu_models, synth):
    """code/train.py:101-112 against the host mirror, call for call: ae.encode(x, is_training=True),
    ae.decode(enc.qbar, True), pc.bitcost(qbar, symbols, True, pad) (batch-statistics batch norm), get_loss with the
    regularisation terms taken from ae / pc as the reference does (:321-326), pc.variables()."""
    from imgcomp_cvpr_b200 import autoencoder, bits, probclass, train, weights as wm
    from oracle import train_oracle as T
    a, p, W = synth('cvpr/low')
    x = wm.synthetic_images(2, 64, 64, seed=22)
    ref = T.training_step(x, W, a, p, dtype=torch.float64, training=True)
    for mode in ('fp32', 'exact'):
        ae = autoencoder.get_network_cls(a)(a, weights=W, mode=mode)
        pc = probclass.get_network_cls(p)(p, num_centers=a.num_centers, weights=W)
        xt = _cuda(x).float()
        enc_out_train = ae.encode(xt, is_training=True)
        x_out_train = ae.decode(enc_out_train.qbar, is_training=True)
        bc_train = pc.bitcost(enc_out_train.qbar, enc_out_train.symbols, is_training=True, pad_value=pc.auto_pad_value(ae))
        bpp_train = bits.bitcost_to_bpp(bc_train, xt).item()
        sym = enc_out_train.symbols.cpu().numpy()
        assert np.array_equal(sym, ref['tensors']['symbols']), mode
        np.testing.assert_allclose(enc_out_train.heatmap.cpu().numpy(), ref['tensors']['heatmap'], atol=5e-4)
        np.testing.assert_allclose(enc_out_train.qbar.cpu().numpy(), ref['tensors']['qbar'], atol=1e-5)
        np.testing.assert_allclose(x_out_train.cpu().numpy(), ref['tensors']['x_out'], atol=3e-2)
        np.testing.assert_allclose(bc_train.cpu().numpy(), ref['tensors']['bc'], atol=3e-3)
        assert abs(bpp_train - ref['tensors']['bc'].sum() / (2 * 64 * 64)) < 1e-4
        # the inference path of the same object is untouched by the training-mode calls
        e_inf = ae.encode(_cuda(x), is_training=False)
        assert e_inf.symbols.shape == enc_out_train.symbols.shape
        a_mse = type(a)(**dict(a.__dict__, distortion_to_minimize='mse'))
        d_train = train.Distortions(a_mse, xt, x_out_train, is_training=True)
        total, H_real, pc_comps, ae_comps = train.get_loss(a, ae, pc, d_train.d_loss_scaled, bc_train, enc_out_train.heatmap)
        assert abs(H_real - ref['H_real']) < 1e-4 and abs(dict(pc_comps)['H_mask'] - ref['H_mask']) < 1e-4
        assert abs(dict(pc_comps)['pc_loss'] - ref['pc_loss']) < 1e-4 * max(1.0, ref['pc_loss'])
        assert abs(dict(ae_comps)['reg_enc_dec'] + dict(pc_comps)['reg'] - ref['reg']) < 1e-6 * ref['reg']
    reg = train.regularization_losses(a, p, W)
    assert abs(ae.encoder_regularization_loss() - reg[0]) < 1e-9 and abs(ae.decoder_regularization_loss() - reg[1]) < 1e-9
    assert pc.regularization_loss() is None                                   # pc_configs/base: regularization_factor = None
    names = [v.name for v in ae.encoder_variables()]
    assert 'autoencoder/encoder/centers' in names and not any('moving_' in n for n in names)      # trainable only
    assert len(names) == 3 * 35 + 1 and len(ae.decoder_variables()) == 3 * 35
    assert [v.name for v in pc.variables()] == pc.variable_names() and len(pc.variables()) == 8
