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
