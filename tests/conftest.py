import os
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


GOLDEN_CASES = {
    # name: (ae config, has real_bpp)
    'tiny_low_1x64x64': ('cvpr/low', True),
    'ragged_low_2x48x72': ('cvpr/low', False),
    'cfg1_low_1x128x128': ('cvpr/low', False),
    'tiny_hi_1x40x24': ('cvpr/hi', False),
}


@pytest.fixture(scope='session')
def synth():
    """name -> (ae_cfg, pc_cfg, weights dict), cached."""
    from imgcomp_cvpr_b200 import config as cfgmod, weights as wm
    cache = {}

    def get(ae_name, pc_name='cvpr/res_shallow', seed=0):
        key = (ae_name, pc_name, seed)
        if key not in cache:
            a, p = cfgmod.ae_config(ae_name), cfgmod.pc_config(pc_name)
            cache[key] = (a, p, wm.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k,
                                                     a.arch_param_B, seed=seed))
        return cache[key]
    return get


def symbol_margin(z64, centers):
    """Distance of the float64 latent to the nearest decision boundary between
    two centres (the quantity that decides whether a symbol may legitimately
    flip between two float32 implementations)."""
    c = np.sort(centers.astype(np.float64))
    mids = (c[1:] + c[:-1]) / 2
    return np.abs(z64[..., None] - mids).min(axis=-1)


@pytest.fixture(scope='session')
def gpu_models(synth):
    """(ae_name, mode) -> (ae, pc, weights) on cuda:0, cached."""
    import torch
    from imgcomp_cvpr_b200 import autoencoder, probclass
    cache = {}

    def get(ae_name='cvpr/low', mode='fp32', pc_name='cvpr/res_shallow'):
        key = (ae_name, mode, pc_name)
        if key not in cache:
            a, p, W = synth(ae_name, pc_name)
            ae = autoencoder.get_network_cls(a)(a, weights=W, mode=mode)
            pc = probclass.get_network_cls(p)(p, num_centers=a.num_centers, weights=W)
            cache[key] = (ae, pc, W)
        return cache[key]
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return get


def golden_projection(name, size):
    """the seeded vector tests/golden/make_train_golden.py projects the gradient of variable `name` on"""
    seed = int.from_bytes(name.encode(), 'little') % (2 ** 31 - 1)
    return np.random.RandomState(seed).standard_normal(size)


def check_training_against_golden(losses, grads, golden, weights, ae_cfg, add_l2, loss_rtol, grad_rtol):
    """A training step (loss components `losses`, gradients `grads` by TF variable name) against the reference-run golden
    of tests/golden/make_train_golden.py.  add_l2: `grads` are without the l2 regularisation terms (the CUDA step adds
    them inside the Adam kernel), add regularization_factor * w / regularization_factor_centers * centers first.
    -> worst relative deviation of a gradient (norm-wise, estimated from norm and projection)."""
    for k in ('total_loss', 'd_loss_scaled', 'pc_loss', 'H_real', 'H_mask', 'ms_ssim'):
        ref = float(golden['loss/' + k])
        assert abs(losses[k] - ref) <= loss_rtol * max(1.0, abs(ref)), (k, losses[k], ref)
    names = [str(n) for n in golden['names']]
    assert sorted(grads) == names
    worst = 0.0
    for i, name in enumerate(names):
        g = np.asarray(grads[name], np.float64)
        if add_l2:
            w = np.asarray(weights[name], np.float64)
            if name.startswith('autoencoder/') and name.endswith('/weights'):
                g = g + ae_cfg.regularization_factor * w
            elif name.endswith('/centers'):
                g = g + ae_cfg.regularization_factor_centers * w
        nref, pref = float(golden['grad_norm'][i]), float(golden['grad_proj'][i])
        dn = abs(np.linalg.norm(g) - nref) / max(nref, 1e-300)
        # the projection on a unit-variance random vector deviates by ~|g - g_ref| (norm-wise): scale by the norm
        dp = abs(float(np.dot(g.ravel(), golden_projection(name, g.size))) - pref) / max(nref, 1e-300)
        assert dn <= grad_rtol and dp <= 4 * grad_rtol, (name, dn, dp)
        worst = max(worst, dn, dp / 4)
        if 'grad/' + name in golden:
            ref = golden['grad/' + name]
            assert np.linalg.norm(g - ref) <= grad_rtol * max(np.linalg.norm(ref), 1e-300), name
    return worst
