"""Host mirror of code/ms_ssim.py (float32 training-loss MS-SSIM)."""
import torch

from . import _lib

_DEFAULTS = dict(max_val=255, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03)


def MultiScaleSSIM(img1, img2, max_val=255, filter_size=11, filter_sigma=1.5,
                   k1=0.01, k2=0.03, weights=None, data_format='NHWC', name=None):
    """MS-SSIM score between img1 and img2: ONE scalar for the batch
    (code/ms_ssim.py:115-186).  Raises RuntimeError like the reference for
    unequal shapes / rank != 4 (code/ms_ssim.py:149-157)."""
    if tuple(img1.shape) != tuple(img2.shape):
        raise RuntimeError('Input images must have the same shape (%s vs. %s).' % (tuple(img1.shape), tuple(img2.shape)))
    if img1.dim() != 4:
        raise RuntimeError('Input images must have four dimensions, not %d' % img1.dim())
    given = dict(max_val=max_val, filter_size=filter_size, filter_sigma=filter_sigma, k1=k1, k2=k2)
    if given != _DEFAULTS or weights is not None:
        raise NotImplementedError('only the default MS-SSIM constants (the ones train.py:431 uses) are built')
    if data_format == 'NHWC':
        img1, img2 = img1.permute(0, 3, 1, 2), img2.permute(0, 3, 1, 2)
    elif data_format != 'NCHW':
        raise ValueError(data_format)
    if img1.shape[1] != 3:
        raise ValueError('expected 3 channels')
    _lib.require_device()
    a, b = img1.contiguous().float(), img2.contiguous().float()
    N, _, H, W = a.shape
    L = _lib.lib()
    ws = torch.empty(L.ic_msssim_workspace_bytes(N, H, W, 0), dtype=torch.uint8, device=a.device)
    out = torch.empty(11, dtype=torch.float32, device=a.device)
    rc = L.ic_msssim_tf_fwd(_lib.ptr(a), _lib.ptr(b), N, H, W, _lib.ptr(out), _lib.c_void_p(out.data_ptr() + 4),
                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    if rc == -1:       # the reference's graph construction fails here (conv kernel larger than the level)
        raise RuntimeError(L.ic_last_error().decode())
    _lib.check(rc)
    MultiScaleSSIM.last_levels = out[1:]
    return out[0]
