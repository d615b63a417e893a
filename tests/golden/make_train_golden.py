"""Golden vectors for the TRAINING step, produced by executing the reference's own graph-construction code.

Runs the UNMODIFIED reference modules autoencoder.py / quantizer.py / probclass.py / ms_ssim.py / bits.py and the
get_loss / Distortions definitions of train.py with is_training=True on tests/tf1_shim/autograd.py (a torch-autograd
stand-in for the TF-1.4 symbols they use; TF itself cannot be installed here), in float64, and differentiates the
total loss like tf.gradients does for get_train_op (code/train.py:86-132,303-349).  Needs /root/reference: run it in
the build container; the .npz it writes travels with the repo.

    python tests/golden/make_train_golden.py

Stored per case: loss components, batch-norm batch statistics of a few layers, and for EVERY variable the gradient's
L2 norm and its projection on a seeded random vector (the full gradients are 9.5 M numbers); small variables in full.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from imgcomp_cvpr_b200 import config, weights          # noqa: E402
from tf1_shim import autograd as shim                 # noqa: E402

CASES = [('train_low_2x64x64', 'cvpr/low', 2, 64, 64, 22), ('train_hi_2x80x48', 'cvpr/hi', 2, 80, 48, 21)]
FULL = ('autoencoder/encoder/centers', 'autoencoder/encoder/h1/weights', 'autoencoder/encoder/to_bn/BatchNorm/gamma',
        'autoencoder/encoder/to_bn/BatchNorm/beta', 'autoencoder/decoder/h13/weights', 'autoencoder/decoder/h13/BatchNorm/beta',
        'probclass3d/logits/conv3d_conv0_mask/weights', 'probclass3d/logits/conv3d_conv0_mask/biases',
        'probclass3d/logits/conv3d_conv2_mask/weights', 'probclass3d/logits/conv3d_conv2_mask/biases',
        'probclass3d/logits/res1/conv3d_conv1_mask/biases')
BN_SCOPES = ('autoencoder/encoder/h1', 'autoencoder/encoder/to_bn', 'autoencoder/decoder/h13',
             'autoencoder/encoder/res_block_enc_2/enc_2_2/conv1', 'autoencoder/decoder/dec_after_res/conv2')


def projection_vector(name, size):
    """seeded per variable name: the same vector is rebuilt by the tests"""
    seed = int.from_bytes(name.encode(), 'little') % (2 ** 31 - 1)
    return np.random.RandomState(seed).standard_normal(size)


def run_reference_training_graph(x_u8, W, ae_cfg, pc_cfg):
    tf = shim.install(W, dtype=torch.float64)
    try:
        import autoencoder
        import probclass
        get_loss, Distortions = shim.load_train_definitions(tf)
        ae = autoencoder.get_network_cls(ae_cfg)(ae_cfg)
        pc = probclass.get_network_cls(pc_cfg)(pc_cfg, num_centers=ae_cfg.num_centers)
        x = shim.AT(torch.tensor(x_u8.astype(np.float64)))
        # code/train.py:101-112
        enc_out = ae.encode(x, is_training=True)
        x_out = ae.decode(enc_out.qbar, is_training=True)
        pc_in = tf.stop_gradient(enc_out.qbar)
        bc = pc.bitcost(pc_in, enc_out.symbols, is_training=True, pad_value=pc.auto_pad_value(ae))
        import bits
        bpp = bits.bitcost_to_bpp(bc, x)
        d = Distortions(ae_cfg, x, x_out, is_training=True)
        total_loss, H_real, pc_comps, ae_comps = get_loss(ae_cfg, ae, pc, d.d_loss_scaled, bc, enc_out.heatmap)
        total_loss.t.backward()
        st = shim.state()
        grads = {k: (v.t.grad.numpy().copy() if v.t.grad is not None else np.zeros(tuple(v.t.shape)))
                 for k, v in st.vars.items() if not (k.endswith('moving_mean') or k.endswith('moving_variance'))}
        comps = dict(pc_comps + ae_comps)
        out = dict(total_loss=float(total_loss), H_real=float(H_real), H_mask=float(comps['H_mask']),
                   pc_loss=float(comps['pc_loss']), d_loss_scaled=float(comps['d_loss_scaled']),
                   reg_enc_dec=float(comps['reg_enc_dec']), ms_ssim=float(d.ms_ssim), bpp=float(bpp), mse=float(d.mse), psnr=float(d.psnr))
        tensors = dict(symbols=enc_out.symbols.t.numpy().copy(), x_out=x_out.t.detach().numpy().copy(),
                       bc=bc.t.detach().numpy().copy(), heatmap=enc_out.heatmap.t.detach().numpy().copy())
        return out, grads, dict(st.bn_stats), tensors
    finally:
        shim.uninstall()


MSSSIM_SHAPES = [(2, 64, 64), (1, 70, 50), (3, 160, 160), (1, 47, 33)]      # the shapes of tests/test_gpu_training_step.py


def msssim_pair(shape, seed=5):
    N, H, W_ = shape
    rng = np.random.RandomState(seed)
    a = rng.uniform(0, 255, size=(N, 3, H, W_)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 12, size=a.shape), 0, 255).astype(np.float32)
    return a, b


def run_reference_msssim(a, b):
    """value and d value / d img2 of the reference's ms_ssim.MultiScaleSSIM (code/ms_ssim.py:115-186), float64"""
    shim.install({}, dtype=torch.float64)
    try:
        import ms_ssim
        bt = torch.tensor(b.astype(np.float64), requires_grad=True)
        v = ms_ssim.MultiScaleSSIM(shim.AT(torch.tensor(a.astype(np.float64))), shim.AT(bt), data_format='NCHW')
        v.t.backward()
        return float(v.t.detach()), bt.grad.numpy().copy()
    finally:
        shim.uninstall()


def main():
    blob = {}
    for shape in MSSSIM_SHAPES:
        a, b = msssim_pair(shape)
        v, g = run_reference_msssim(a, b)
        key = 'x'.join(str(s) for s in shape)
        blob['value/' + key] = np.float64(v)
        blob['grad_norm/' + key] = np.float64(np.linalg.norm(g))
        blob['grad_proj/' + key] = np.float64(np.dot(g.ravel(), projection_vector('msssim' + key, g.size)))
        blob['grad_corner/' + key] = g[0, :, :6, :6].copy()
        blob['grad_tail/' + key] = g[-1, :, -6:, -6:].copy()
        print('ms-ssim', shape, v, np.linalg.norm(g))
    np.savez_compressed(os.path.join(HERE, 'train_msssim_grad.npz'), **blob)
    for tag, ae_name, N, H, W_, seed in CASES:
        a, p = config.ae_config(ae_name), config.pc_config('cvpr/res_shallow')
        Wt = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
        x = weights.synthetic_images(N, H, W_, seed=seed)
        out, grads, bn, tensors = run_reference_training_graph(x, Wt, a, p)
        blob = {'loss/' + k: np.float64(v) for k, v in out.items()}
        names = sorted(grads)
        blob['names'] = np.array(names)
        blob['grad_norm'] = np.array([np.linalg.norm(grads[k]) for k in names])
        blob['grad_proj'] = np.array([float(np.dot(grads[k].ravel(), projection_vector(k, grads[k].size))) for k in names])
        for k in FULL:
            blob['grad/' + k] = grads[k]
        for s in BN_SCOPES:
            blob['bn_mean/' + s], blob['bn_var_unbiased/' + s] = bn[s]
        blob['symbols'] = tensors['symbols'].astype(np.uint8)
        blob['bc_sum_per_image'] = tensors['bc'].sum(axis=(1, 2, 3))
        blob['x_out_mean_per_image'] = tensors['x_out'].mean(axis=(1, 2, 3))
        blob['heatmap_sum'] = np.float64(tensors['heatmap'].sum())
        blob['meta'] = np.array([ae_name, str(N), str(H), str(W_), str(seed)])
        path = os.path.join(HERE, tag + '.npz')
        np.savez_compressed(path, **blob)
        print(tag, {k: round(v, 6) for k, v in out.items()}, '%d variables' % len(names), '%.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
