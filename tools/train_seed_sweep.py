"""Gradient error of the training step (GPU float32 vs float64 oracle) over a few image seeds: shows that the
typical norm-wise error is ~1e-5 and that the occasional 1e-2 layer is a discrete ReLU / clip flip."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imgcomp_cvpr_b200 import config, trainer, weights
from oracle import train_oracle as T

a, p = config.ae_config('cvpr/low'), config.pc_config('cvpr/res_shallow')
W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
tr = trainer.Trainer(a, p, W)
for seed in range(21, 29):
    x = weights.synthetic_images(2, 64, 64, seed=seed)
    ref = T.training_step(x, W, a, p, dtype=torch.float64, training=True)
    out = tr.forward_backward(torch.from_numpy(x).cuda(), is_training=True, update_moving=False)
    G = tr.gradients()
    mism = int((out['tensors']['symbols'].cpu().numpy() != ref['tensors']['symbols']).sum())
    errs = []
    for name, g_ref in ref['grads'].items():
        g = G[name].astype(np.float64)
        if name.startswith('autoencoder/') and name.endswith('/weights'):
            g = g + a.regularization_factor * np.asarray(W[name], np.float64)
        elif name.endswith('/centers'):
            g = g + a.regularization_factor_centers * np.asarray(W[name], np.float64)
        errs.append((float(np.linalg.norm(g - g_ref) / max(np.linalg.norm(g_ref), 1e-30)), name))
    errs.sort(reverse=True)
    print('seed %d: symbol mismatches %d, max err %.2e (%s), median %.2e, > 1e-3: %d' %
          (seed, mism, errs[0][0], errs[0][1].split('autoencoder/')[-1], errs[len(errs) // 2][0], sum(e > 1e-3 for e, _ in errs)), flush=True)
