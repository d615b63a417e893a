"""GPU debugging aid for the tcgen05 conv: one layer, TC vs FFMA, with error maps.
usage: python tools/tc_debug.py N H W mode [res]"""
import sys
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
from imgcomp_cvpr_b200 import _lib, autoencoder, config, weights


def run(ae, layer, x, r1, r2, mode):
    L = _lib.lib()
    N, H, W, _ = x.shape
    out = torch.full_like(x, float('nan'))
    ws = torch.empty(4 * x.numel() * 4 + 4096, dtype=torch.uint8, device='cuda')
    _lib.check(L.ic_debug_conv3x3(ae._handle, 0, layer, _lib.ptr(x), _lib.ptr(r1), _lib.ptr(r2), N, H, W, _lib.ptr(out),
                                  _lib.ptr(ws), ws.numel(), _lib.MODES[mode], _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out


def main():
    N, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    mode = sys.argv[4]
    use_res = len(sys.argv) > 5
    layer = int(os.environ.get('LAYER', '0'))
    a = config.ae_config('cvpr/low')
    Wt = weights.synthetic_weights()
    ae = autoencoder.get_network_cls(a)(a, weights=Wt)
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn((N, H, W, 128), device='cuda', generator=g)
    r1 = torch.randn((N, H, W, 128), device='cuda', generator=g) if use_res else None
    r2 = torch.randn((N, H, W, 128), device='cuda', generator=g) if use_res else None
    ref = run(ae, layer, x, r1, r2, 'fp32')
    out = run(ae, layer, x, r1, r2, mode)
    err = (out - ref).abs()
    nan = torch.isnan(out).float().mean().item()
    print('N,H,W=%d,%d,%d mode=%s res=%s layer=%d: max err %.3e mean err %.3e ref rms %.3f nan frac %.3f' % (
        N, H, W, mode, use_res, layer, torch.nan_to_num(err, 1e9).max().item(), torch.nan_to_num(err, 0).mean().item(),
        ref.pow(2).mean().sqrt().item(), nan))
    tol = 2e-5 if mode == 'exact' else 3e-2
    if torch.nan_to_num(err, 1e9).max().item() > tol * max(1.0, ref.abs().max().item()):
        e = torch.nan_to_num(err, 9.0)
        print('err by y (max):', ['%.1e' % v for v in e.amax(dim=(0, 2, 3)).tolist()][:40])
        print('err by x (max):', ['%.1e' % v for v in e.amax(dim=(0, 1, 3)).tolist()][:40])
        print('err by chunk (max):', ['%.1e' % v for v in e.reshape(N, H, W, 16, 8).amax(dim=(0, 1, 2, 4)).tolist()])
        print('sample out', out[0, 0, 0, :8].tolist())
        print('sample ref', ref[0, 0, 0, :8].tolist())
        print('FAIL')
        sys.exit(1)
    print('OK')


if __name__ == '__main__':
    main()
