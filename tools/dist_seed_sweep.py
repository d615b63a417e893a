"""mse / psnr distortion training step against the float64 oracle over image seeds: a clip of x_out at 0 / 255 or a ReLU
input landing on the other side of zero in float32 than in float64 is a discrete event (like a symbol flip); the test uses a
seed without one."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
from imgcomp_cvpr_b200 import config as cfgmod, trainer, weights
from oracle import train_oracle as T


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


a0, p = cfgmod.ae_config('cvpr/low'), cfgmod.pc_config('cvpr/res_shallow')
Wt = weights.synthetic_weights(a0.num_chan_bn, a0.num_centers, p.arch_param__k, a0.arch_param_B, seed=0)
for kind in ('mse', 'psnr'):
    a = cfgmod.Config(**dict(vars(a0), distortion_to_minimize=kind))
    for seed in range(20, 30):
        x = weights.synthetic_images(2, 48, 40, seed=seed)
        tr = trainer.Trainer(a, p, Wt, num_itr_per_epoch=100, mode='fp32')
        ref = T.training_step(x, Wt, a, p, dtype=torch.float64, training=True)
        out = tr.forward_backward(torch.from_numpy(x).cuda(), is_training=True, update_moving=False)
        mism = int((out['tensors']['symbols'].cpu().numpy() != ref['tensors']['symbols']).sum())
        G = tr.gradients()
        errs = []
        for name, g_ref in ref['grads'].items():
            g = G[name].astype(np.float64)
            w = np.asarray(Wt[name], np.float64)
            if name.startswith('autoencoder/') and name.endswith('/weights'):
                g = g + a.regularization_factor * w
            elif name.endswith('/centers'):
                g = g + a.regularization_factor_centers * w
            errs.append((rel(g, g_ref), name))
        errs.sort(reverse=True)
        print('%s seed %d: symbol mismatches %d, d_loss %.4f / %.4f, worst grad %.2e (%s), median %.2e' % (
            kind, seed, mism, out['d_loss_scaled'], ref['d_loss_scaled'], errs[0][0], errs[0][1][-40:], errs[len(errs) // 2][0]), flush=True)
