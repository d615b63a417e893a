"""cfg 3 (BASELINE.json configs[2]): the training step -- forward + MS-SSIM / rate loss + backward + two Adam groups --
of imgcomp_cvpr_b200.trainer against the CPU oracle (oracle/train_oracle.py: the reference graph of
code/train.py:86-132,303-349 restated on torch-CPU float64 with autograd doing the differentiation).
Tolerances: float32 kernels vs float64 oracle through ~70 conv+BN layers -> norm-wise 5e-3 on gradients,
1e-4 relative on the loss scalars (written next to each assert)."""
import numpy as np
import pytest
import torch

from oracle import train_oracle as T

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize('shape', [(2, 64, 64), (1, 70, 50), (3, 160, 160), (1, 47, 33)])
def test_msssim_backward(shape):
    """d MS-SSIM / d img2 (ic_msssim_tf_bwd) vs autograd of the oracle's ms_ssim_tf; odd sizes exercise the REFLECT
    padding of both the downsample and the small-level blur (code/ms_ssim.py:24-29,46-64)."""
    from imgcomp_cvpr_b200 import nn
    N, H, W = shape
    rng = np.random.RandomState(5)
    a = rng.uniform(0, 255, size=(N, 3, H, W)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 12, size=a.shape), 0, 255).astype(np.float32)
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    val = T.ms_ssim_tf(torch.tensor(a, dtype=torch.float64), bt)
    (-5000.0 * val).backward()
    d, v = nn.msssim_tf_bwd(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), -5000.0)
    assert abs(float(v.item()) - float(val.detach())) < 2e-5
    err = _rel(d.cpu().numpy(), bt.grad.numpy())
    print('ms-ssim bwd %s: value %.6f, rel grad err %.2e' % (shape, float(val), err))
    assert err < 2e-3


def test_msssim_backward_rejects_what_the_reference_rejects():
    """a level whose height is below the tap count after the W-derived padding: the reference's graph construction fails
    (code/ms_ssim.py:24-29), the library returns IC_ERR_INVALID"""
    from imgcomp_cvpr_b200 import _lib, nn
    a = torch.zeros((1, 3, 50, 70), device='cuda')
    with pytest.raises(_lib.IcError):
        nn.msssim_tf_bwd(a, a.clone(), 1.0)


def _setup(synth, ae_name, N, H, W, seed=21, mode='fp32'):
    from imgcomp_cvpr_b200 import trainer, weights
    ae_cfg, pc_cfg, Wt = synth(ae_name)
    x = weights.synthetic_images(N, H, W, seed=seed)
    tr = trainer.Trainer(ae_cfg, pc_cfg, Wt, num_itr_per_epoch=100, mode=mode)
    return ae_cfg, pc_cfg, Wt, x, tr


@pytest.mark.parametrize('N,H,W', [(2, 16, 16), (3, 40, 40), (1, 19, 27)])
def test_conv3x3_tensor_core_forward_and_data_gradient(N, H, W):
    """ic_nn_conv3x3_tc (tcgen05, fp16 hi/lo, weights re-packed on the device) against the float32 FFMA primitives it
    replaces in the training step; tolerance 2e-5 of the output range (measured ~1e-6: DESIGN.md 4.2)."""
    from imgcomp_cvpr_b200 import nn
    rng = np.random.RandomState(3)
    x = torch.from_numpy(rng.randn(N, H, W, 128).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.randn(3, 3, 128, 128) * 0.05).astype(np.float32)).cuda()
    y_ref, y = nn.conv2d_fwd(x, w), nn.conv3x3_tc(x, w)
    assert float((y - y_ref).abs().max()) <= 2e-5 * float(y_ref.abs().max())
    dy = torch.from_numpy(rng.randn(N, H, W, 128).astype(np.float32)).cuda()
    dx_ref, dx = nn.conv2d_bwd_data(dy, w, x.shape), nn.conv3x3_tc(dy, w, data_grad=True)
    assert float((dx - dx_ref).abs().max()) <= 2e-5 * float(dx_ref.abs().max())
    # a second call with other weights: the device-side re-pack really follows the weights
    w2 = w * 3.0 + 0.01
    y2_ref, y2 = nn.conv2d_fwd(x, w2), nn.conv3x3_tc(x, w2)
    assert float((y2 - y2_ref).abs().max()) <= 2e-5 * float(y2_ref.abs().max())


@pytest.mark.parametrize('N,H,W,gscale', [(2, 16, 16, 1.0), (3, 40, 40, 1e-5), (1, 19, 27, 300.0), (4, 8, 8, 1.0)])
def test_conv3x3_tensor_core_backward(N, H, W, gscale):
    """ic_nn_conv3x3_tc_bwd: data gradient (conv_tc with flipped, transposed taps) and filter gradient (tcgen05 GEMM
    over pixels on MN-major operands, split over pixel tiles) against the float32 FFMA primitives; gscale exercises the
    per-tensor power-of-two pre-scaling (gradients of 1e-5 must keep float32-class precision in the fp16 hi/lo split)."""
    from imgcomp_cvpr_b200 import nn
    rng = np.random.RandomState(4)
    x = torch.from_numpy(np.maximum(rng.randn(N, H, W, 128), 0).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.randn(3, 3, 128, 128) * 0.05).astype(np.float32)).cuda()
    dy = torch.from_numpy((rng.randn(N, H, W, 128) * gscale).astype(np.float32)).cuda()
    dx_ref = nn.conv2d_bwd_data(dy, w, x.shape)
    dw_ref = nn.conv2d_bwd_filter(x, dy, w.shape)
    dx, dw = nn.conv3x3_tc_bwd(x, dy, w)
    e_dx = float((dx - dx_ref).abs().max()) / float(dx_ref.abs().max())
    e_dw = float((dw - dw_ref).abs().max()) / float(dw_ref.abs().max())
    print('conv3x3 tc bwd %s gscale %g: max err dx %.2e dw %.2e (of the range)' % ((N, H, W), gscale, e_dx, e_dw))
    assert e_dx <= 2e-5 and e_dw <= 2e-5
    # float64 truth for the filter gradient on the small case
    if N * H * W <= 1024:
        xt = torch.tensor(x.cpu().numpy().transpose(0, 3, 1, 2), dtype=torch.float64)
        wt = torch.tensor(w.cpu().numpy(), dtype=torch.float64, requires_grad=True)
        y = T.conv2d_same(xt, wt, 1)
        (y * torch.tensor(dy.cpu().numpy().transpose(0, 3, 1, 2), dtype=torch.float64)).sum().backward()
        ref = wt.grad.numpy()
        assert np.abs(dw.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    _, dw_only = nn.conv3x3_tc_bwd(x, dy, w, need_dx=False)
    assert torch.equal(dw_only, dw)
    # the forward pass kept for the backward pass (input planes + scales): no second maximum search / split, same bits
    y_keep, cache = nn.conv3x3_tc(x, w, keep=True)
    assert torch.equal(y_keep, nn.conv3x3_tc(x, w))
    dx_c, dw_c = nn.conv3x3_tc_bwd(x, dy, w, cache=cache)
    assert torch.equal(dx_c, dx) and torch.equal(dw_c, dw)


def test_exact_mode_step_matches_fp32_mode(synth):
    """Trainer(mode='exact') (3x3 convs' forward / data gradient on tensor cores) against Trainer(mode='fp32')"""
    ae_cfg, pc_cfg, Wt, x, tr32 = _setup(synth, 'cvpr/low', 2, 64, 64, seed=22)
    _, _, _, _, trx = _setup(synth, 'cvpr/low', 2, 64, 64, seed=22, mode='exact')
    xg = torch.from_numpy(x).cuda()
    a = tr32.forward_backward(xg)
    b = trx.forward_backward(xg)
    mism = int((a['tensors']['symbols'] != b['tensors']['symbols']).sum())
    print('exact vs fp32 training forward: symbol mismatches %d, loss %.4f vs %.4f' % (mism, b['total_loss'], a['total_loss']))
    assert mism <= 2
    if mism == 0:
        assert abs(a['total_loss'] - b['total_loss']) <= 2e-4 * abs(a['total_loss'])
        Ga, Gb = tr32.gradients(), trx.gradients()
        errs = sorted((_rel(Gb[k], Ga[k]), k) for k in Ga)
        print('  worst gradient difference %.2e (%s), median %.2e' % (errs[-1][0], errs[-1][1], errs[len(errs) // 2][0]))
        assert errs[len(errs) // 2][0] < 2e-4


@pytest.mark.parametrize('ae_name,N,H,W,seed', [('cvpr/low', 2, 64, 64, 22), ('cvpr/hi', 2, 80, 48, 21)])
def test_training_step_matches_oracle(synth, ae_name, N, H, W, seed):
    ae_cfg, pc_cfg, Wt, x, tr = _setup(synth, ae_name, N, H, W, seed=seed)
    ref = T.training_step(x, Wt, ae_cfg, pc_cfg, dtype=torch.float64, training=True)
    out = tr.forward_backward(torch.from_numpy(x).cuda(), is_training=True, update_moving=False)
    sym = out['tensors']['symbols'].cpu().numpy()
    mism = int((sym != ref['tensors']['symbols']).sum())
    print('%s: symbol mismatches %d / %d' % (ae_name, mism, sym.size))
    assert mism == 0, 'a symbol flipped between the float32 kernels and the float64 oracle: pick another seed'
    for k in ('total_loss', 'd_loss_scaled', 'pc_loss', 'H_real', 'H_mask', 'ms_ssim', 'reg'):
        print('  %-14s gpu %.6f  oracle %.6f' % (k, out[k], ref[k]))
        assert abs(out[k] - ref[k]) <= 1e-4 * max(1.0, abs(ref[k])), k           # 1e-4 relative (north star: bpp / MS-SSIM 1e-4)
    np.testing.assert_allclose(out['tensors']['x_out'].cpu().numpy(), ref['tensors']['x_out'], atol=2e-2)
    np.testing.assert_allclose(out['tensors']['bc'].cpu().numpy(), ref['tensors']['bc'], atol=2e-4, rtol=1e-4)
    G = tr.gradients()
    errs = []
    f = ae_cfg.regularization_factor
    for name, g_ref in ref['grads'].items():
        g = G[name].astype(np.float64)
        w = np.asarray(Wt[name], np.float64)
        if name.startswith('autoencoder/') and name.endswith('/weights'):
            g = g + f * w                                    # the oracle differentiates the l2 terms too
        elif name.endswith('/centers'):
            g = g + ae_cfg.regularization_factor_centers * w
        assert g.shape == g_ref.shape, name
        if name.startswith('probclass3d') and name.endswith('/weights'):
            # masked taps get no gradient (code/probclass.py:252-253)
            assert np.abs(g[1, 2]).max() == 0 and np.abs(g[1, 1, 2]).max() == 0, name
        e = _rel(g, g_ref)
        errs.append((e, name))
    errs.sort(reverse=True)
    for e, name in errs[:8]:
        print('  grad err %.2e  %s' % (e, name))
    print('  median gradient error %.2e over %d variables' % (errs[len(errs) // 2][0], len(errs)))
    # float32 kernels against the float64 oracle: norm-wise 1e-3 on every variable, 1e-4 in the median (measured:
    # worst 6e-5, median 6e-6).  The image seeds are ones where no ReLU input / heatmap clip lands on the other side
    # of zero in float32 than in float64; such a flip is a discrete event that moves one layer's gradient by ~1e-2 and
    # everything upstream of it by ~5e-3 (tools/train_seed_sweep.py: seeds 21, 24 of cvpr/low), like a symbol flip.
    assert errs[0][0] < 1e-3, errs[:8]
    assert errs[len(errs) // 2][0] < 1e-4


@pytest.mark.parametrize('tag,ae_name,N,H,W,seed', [('train_low_2x64x64', 'cvpr/low', 2, 64, 64, 22),
                                                     ('train_hi_2x80x48', 'cvpr/hi', 2, 80, 48, 21)])
def test_training_step_matches_reference_run_golden(synth, tag, ae_name, N, H, W, seed):
    """The CUDA step against the vectors tests/golden/make_train_golden.py produced by executing the reference's own
    training graph (unmodified modules on the autograd TF1 shim, float64): symbols identical, loss components within
    1e-4 relative, every variable's gradient within 1e-3 norm-wise (norm and seeded projection; small variables in full)."""
    from conftest import check_training_against_golden, load_golden
    g = load_golden(tag)
    ae_cfg, pc_cfg, Wt, x, tr = _setup(synth, ae_name, N, H, W, seed=seed)
    out = tr.forward_backward(torch.from_numpy(x).cuda(), is_training=True, update_moving=False)
    assert np.array_equal(out['tensors']['symbols'].cpu().numpy().astype(np.uint8), g['symbols'])
    worst = check_training_against_golden(out, tr.gradients(), g, Wt, ae_cfg, True, 1e-4, 1e-3)
    print('%s vs reference-run golden: worst gradient deviation %.2e' % (tag, worst))


def test_step_applies_adam_and_moving_averages(synth):
    """tr.step = forward_backward + tf.train.AdamOptimizer._apply_dense per group (code/train.py:339-349,
    training_helpers.py:22-48) + the decay-0.9 moving averages of slim.batch_norm (autoencoder.py:115-125)."""
    ae_cfg, pc_cfg, Wt, x, tr = _setup(synth, 'cvpr/low', 2, 64, 64)
    xg = torch.from_numpy(x).cuda()
    tr.forward_backward(xg, is_training=True, update_moving=False)
    G = tr.gradients()
    W0 = tr.weights()
    ref = T.training_step(x, Wt, ae_cfg, pc_cfg, dtype=torch.float64, training=True)
    out = tr.step(xg)
    assert tr.global_step == 1
    W1 = tr.weights()
    lr_ae, lr_pc = ae_cfg.lr_initial, pc_cfg.lr_initial
    for name in G:
        if name.startswith('probclass3d'):
            lr, l2 = lr_pc, 0.0
        elif name.endswith('/weights'):
            lr, l2 = lr_ae, ae_cfg.regularization_factor
        elif name.endswith('/centers'):
            lr, l2 = lr_ae, ae_cfg.regularization_factor_centers
        else:
            lr, l2 = lr_ae, 0.0
        g = G[name].astype(np.float64) + l2 * W0[name].astype(np.float64)
        w_ref, _, _ = T.adam_update(W0[name].astype(np.float64), g, 0.0, 0.0, 1, lr)
        # the first Adam step moves every weight with a non-zero gradient by ~lr: compare the UPDATE, not the weight
        upd, upd_ref = W1[name].astype(np.float64) - W0[name], w_ref - W0[name]
        big = np.abs(g) > 1e-6 * np.abs(g).max()
        assert np.abs(upd - upd_ref)[big].max() <= 0.02 * lr + 1e-9, name
    for scope, (mu, unb) in ref['bn_stats'].items():
        mm0, mv0 = Wt[scope + '/BatchNorm/moving_mean'], Wt[scope + '/BatchNorm/moving_variance']
        np.testing.assert_allclose(W1[scope + '/BatchNorm/moving_mean'], 0.9 * mm0 + 0.1 * mu, rtol=2e-3, atol=2e-4 * np.abs(mu).max())
        np.testing.assert_allclose(W1[scope + '/BatchNorm/moving_variance'], 0.9 * mv0 + 0.1 * unb, rtol=2e-3)
    assert np.isfinite(out['total_loss'])


def test_loss_decreases_on_a_fixed_batch(synth):
    from imgcomp_cvpr_b200 import config
    ae_cfg, pc_cfg, Wt, x, tr = _setup(synth, 'cvpr/med', 4, 64, 64)
    xg = torch.from_numpy(x).cuda()
    losses = [tr.step(xg)['total_loss'] for _ in range(12)]
    print('losses', ['%.2f' % l for l in losses])
    assert losses[-1] < losses[0]
    # the trained variables load into the inference classes unchanged
    from imgcomp_cvpr_b200 import autoencoder, probclass
    W1 = tr.weights()
    ae = autoencoder.get_network_cls(ae_cfg)(ae_cfg, weights=W1)
    pc = probclass.get_network_cls(pc_cfg)(pc_cfg, num_centers=ae_cfg.num_centers, weights=W1)
    enc = ae.encode(xg, is_training=False)
    bc = pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
    assert torch.isfinite(bc).all()


def test_inference_mode_matches_pinned_forward(synth):
    """is_training=False runs the same primitives on the moving statistics: must reproduce the inference graph the
    reference-run goldens pin (oracle/imgcomp_oracle.py)."""
    from oracle import imgcomp_oracle as O
    ae_cfg, pc_cfg, Wt, x, tr = _setup(synth, 'cvpr/low', 1, 64, 64, seed=11)
    out = tr.forward_backward(torch.from_numpy(x).cuda(), is_training=False, backward=False)
    ref = O.val_forward(x, Wt, ae_cfg.num_chan_bn)
    assert (out['tensors']['symbols'].cpu().numpy() != ref['enc']['symbols']).sum() == 0
    assert abs(out['bpp'] - ref['bpp'][0]) < 1e-4


def test_cuda_graph_step_is_the_eager_step(synth):
    """enable_cuda_graph(): forward + loss + backward + moving averages + Adam replayed as one CUDA graph must give
    exactly what the eager launches give (every reduction in the step has a fixed order)."""
    ae_cfg, pc_cfg, Wt, x, tr_e = _setup(synth, 'cvpr/low', 2, 64, 64, seed=22, mode='exact')
    _, _, _, _, tr_g = _setup(synth, 'cvpr/low', 2, 64, 64, seed=22, mode='exact')
    from imgcomp_cvpr_b200 import weights
    xs = [torch.from_numpy(weights.synthetic_images(2, 64, 64, seed=30 + i)).cuda() for i in range(3)]
    tr_g.enable_cuda_graph(xs[0])
    for xi in xs:
        a, b = tr_e.step(xi), tr_g.step(xi)
        for k in ('total_loss', 'd_loss_scaled', 'pc_loss', 'H_real', 'H_mask', 'ms_ssim', 'bpp'):
            assert a[k] == b[k], (k, a[k], b[k])
    assert tr_e.global_step == tr_g.global_step == 3
    We, Wg = tr_e.weights(), tr_g.weights()
    for k in We:
        assert np.array_equal(We[k], Wg[k]), k


def test_cuda_graph_survives_growth_of_the_shared_workspace(synth):
    """The captured step holds raw pointers into nn.py's shared scratch buffer; a later, larger call replaces that buffer.
    The graph owner keeps the old block alive, so replays neither read nor corrupt memory handed to other tensors."""
    from imgcomp_cvpr_b200 import nn, weights
    ae_cfg, pc_cfg, Wt, x, tr_e = _setup(synth, 'cvpr/low', 2, 64, 64, seed=22, mode='exact')
    _, _, _, _, tr_g = _setup(synth, 'cvpr/low', 2, 64, 64, seed=22, mode='exact')
    xs = [torch.from_numpy(weights.synthetic_images(2, 64, 64, seed=40 + i)).cuda() for i in range(3)]
    tr_g.enable_cuda_graph(xs[0])
    ws_before = nn.current_workspace()
    big = torch.from_numpy(weights.synthetic_images(4, 160, 160, seed=3)).cuda()
    _, _, _, _, tr_big = _setup(synth, 'cvpr/low', 4, 160, 160, seed=3, mode='exact')
    tr_big.forward_backward(big)                                  # needs a larger scratch buffer -> replaced
    assert nn.current_workspace() is not ws_before and tr_g._graph_ws is ws_before
    # the graph owner's OWN larger eager call: replaces its batch-norm partial-sum buffer (fused trunk), which the captured
    # launches also point into; the graph keeps the captured one
    partial_before = tr_g._bn_partial
    tr_g.forward_backward(big)
    assert tr_g._bn_partial is not partial_before and tr_g._graph_partial is partial_before
    filler = [torch.full((1 << 20,), 7.0, device='cuda') for _ in range(64)]        # whatever the allocator hands out next
    for xi in xs:
        a, b = tr_e.step(xi), tr_g.step(xi)
        for k in ('total_loss', 'd_loss_scaled', 'pc_loss', 'H_real', 'H_mask', 'ms_ssim', 'bpp'):
            assert a[k] == b[k], (k, a[k], b[k])
    assert all(float(f.min()) == 7.0 and float(f.max()) == 7.0 for f in filler)


@pytest.mark.parametrize('minimize_for', ['mse', 'psnr'])
def test_training_step_mse_and_psnr_distortions_match_oracle(synth, minimize_for):
    """config.distortion_to_minimize = 'mse' / 'psnr' (code/train.py:381-397: float32 squared error while training; the
    psnr variant minimises K_psnr - mean_n 10 log10(255^2 / mse_n)): loss components and every gradient against the
    float64 oracle, same gates as the MS-SSIM step."""
    from imgcomp_cvpr_b200 import config as cfgmod, trainer, weights
    ae_cfg, pc_cfg, Wt = synth('cvpr/low')
    ae_cfg = cfgmod.Config(**dict(vars(ae_cfg), distortion_to_minimize=minimize_for))
    # 48 x 40 is below the 11-tap MS-SSIM's minimum size: MS-SSIM is never evaluated.  Seed: tools/dist_seed_sweep.py (7 of 10
    # seeds have every gradient within 3e-5; seeds 22, 25, 29 have a ReLU / clip decision that differs between float32 and
    # float64, a discrete event that moves one layer by ~1e-2 -- see test_training_step_matches_oracle)
    x = weights.synthetic_images(2, 48, 40, seed=23)
    tr = trainer.Trainer(ae_cfg, pc_cfg, Wt, num_itr_per_epoch=100, mode='fp32')
    ref = T.training_step(x, Wt, ae_cfg, pc_cfg, dtype=torch.float64, training=True)
    out = tr.forward_backward(torch.from_numpy(x).cuda(), is_training=True, update_moving=False)
    mism = int((out['tensors']['symbols'].cpu().numpy() != ref['tensors']['symbols']).sum())
    assert mism == 0, 'a symbol flipped between the float32 kernels and the float64 oracle: pick another seed'
    assert out['ms_ssim'] is None
    for k in ('total_loss', 'd_loss_scaled', 'pc_loss', 'H_real', 'H_mask', 'reg'):
        print('  %-14s gpu %.6f  oracle %.6f' % (k, out[k], ref[k]))
        assert abs(out[k] - ref[k]) <= 1e-4 * max(1.0, abs(ref[k])), k
    G = tr.gradients()
    errs = []
    for name, g_ref in ref['grads'].items():
        g = G[name].astype(np.float64)
        w = np.asarray(Wt[name], np.float64)
        if name.startswith('autoencoder/') and name.endswith('/weights'):
            g = g + ae_cfg.regularization_factor * w
        elif name.endswith('/centers'):
            g = g + ae_cfg.regularization_factor_centers * w
        errs.append((_rel(g, g_ref), name))
    errs.sort(reverse=True)
    print('  %s: worst gradient error %.2e (%s), median %.2e' % (minimize_for, errs[0][0], errs[0][1], errs[len(errs) // 2][0]))
    assert errs[0][0] < 1e-3, errs[:8]
    assert errs[len(errs) // 2][0] < 1e-4


def test_distortion_backward_kernel():
    """ic_nn_distortion_bwd against the closed form in float64 (code/train.py:381-397,420-425)"""
    from imgcomp_cvpr_b200 import nn
    rng = np.random.RandomState(9)
    x = rng.uniform(0, 255, size=(3, 3, 24, 40)).astype(np.float32)
    y = np.clip(x + rng.normal(0, 20, size=x.shape), 0, 255).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64)
    for psnr in (False, True):
        yt = torch.tensor(y, dtype=torch.float64, requires_grad=True)
        mse = ((yt - xt) ** 2).mean(dim=(1, 2, 3))
        loss = (100.0 - (10 * torch.log10(255.0 * 255.0 / mse)).mean()) if psnr else mse.mean()
        loss.backward()
        d, m = nn.distortion_bwd(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), psnr=psnr)
        np.testing.assert_allclose(m.cpu().numpy(), mse.detach().numpy(), rtol=1e-6)
        assert _rel(d.cpu().numpy(), yt.grad.numpy()) < 1e-6


def test_fused_trunk_matches_unfused_sequence(synth, monkeypatch):
    """mode='exact' with the trunk fusions (batch norm writes the next conv's fp16 planes, the conv's output pass accumulates
    the next batch norm's statistics, one weight-scale launch per step) against the unfused sequence of the same kernels
    (IC_TRAIN_FUSED=0).  The statistics are summed in the same grouping, so the only difference is that activations are
    split unscaled instead of pre-scaled by a power of two: identical symbols, losses to 1e-6, gradients to 1e-4."""
    outs, grads = [], []
    for fused in ('0', '1'):
        monkeypatch.setenv('IC_TRAIN_FUSED', fused)
        ae_cfg, pc_cfg, Wt, x, tr = _setup(synth, 'cvpr/low', 2, 64, 64, seed=22, mode='exact')
        outs.append(tr.forward_backward(torch.from_numpy(x).cuda()))
        grads.append(tr.gradients())
    monkeypatch.delenv('IC_TRAIN_FUSED', raising=False)
    a, b = outs
    assert torch.equal(a['tensors']['symbols'], b['tensors']['symbols'])
    for k in ('total_loss', 'd_loss_scaled', 'pc_loss', 'ms_ssim'):
        assert abs(a[k] - b[k]) <= 1e-6 * max(1.0, abs(a[k])), (k, a[k], b[k])
    errs = sorted((_rel(grads[1][k], grads[0][k]), k) for k in grads[0])
    print('fused vs unfused trunk: worst gradient difference %.2e (%s), median %.2e' % (errs[-1][0], errs[-1][1], errs[len(errs) // 2][0]))
    assert errs[-1][0] < 1e-4
