"""Is the tensor-core accumulation error a systematic (round-toward-zero) bias?  One layer, fp64 truth."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from imgcomp_cvpr_b200 import _lib, autoencoder, config, weights
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))

def conv(ae, layer, x, mode):
    L = _lib.lib(); N, H, W, _ = x.shape
    out = torch.empty_like(x); ws = torch.empty(4 * x.numel() * 4 + 4096, dtype=torch.uint8, device='cuda')
    _lib.check(L.ic_debug_conv3x3(ae._handle, 0, layer, _lib.ptr(x), None, None, N, H, W, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.MODES[mode], _lib.stream_ptr()))
    return out

a = config.ae_config('cvpr/low'); Wt = weights.synthetic_weights()
ae = autoencoder.get_network_cls(a)(a, weights=Wt)
import json
enc_scopes = weights.conv_scopes(5)[0]
g = torch.Generator(device='cuda').manual_seed(3)
for layer in (1, 11, 31):
    scope = enc_scopes[layer]
    w = torch.from_numpy(Wt[scope + '/weights']).double().cuda().permute(3, 2, 0, 1)
    bn = {k: torch.from_numpy(Wt[scope + '/BatchNorm/' + k]).double().cuda() for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')}
    sc = bn['gamma'] / torch.sqrt(bn['moving_variance'] + 1e-5)
    for name, x in (('randn', torch.randn((1, 64, 64, 128), device='cuda', generator=g)),
                    ('relu', torch.relu(torch.randn((1, 64, 64, 128), device='cuda', generator=g)) * 1.5),
                    ('big', torch.randn((1, 64, 64, 128), device='cuda', generator=g) * 20 + 3)):
        pre = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
        # only layers without ReLU allow backing out the accumulator; conv2 layers (odd index) have none
        out = conv(ae, layer, x, 'exact').double()
        acc = (out - bn['beta']) / sc + bn['moving_mean']
        m = pre.abs() > 0.2 * pre.pow(2).mean().sqrt()
        rel = ((acc - pre) / pre)[m]
        out32 = conv(ae, layer, x, 'fp32').double()
        acc32 = (out32 - bn['beta']) / sc + bn['moving_mean']
        rel32 = ((acc32 - pre) / pre)[m]
        print('layer %2d %-5s: exact mean rel err %.3e (std %.2e)   fp32 mean rel err %.3e (std %.2e)' % (
            layer, name, rel.mean().item(), rel.std().item(), rel32.mean().item(), rel32.std().item()))
