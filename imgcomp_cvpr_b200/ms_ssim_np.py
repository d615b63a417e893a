"""Host mirror of code/ms_ssim_np.py: the float64 MS-SSIM on uint8 that
val.py:93 evaluates per image through tf.py_func."""
import torch

from . import _lib


def MultiScaleSSIM_batch(img1, img2, data_format='NCHW'):
    """float64 MS-SSIM PER IMAGE (the reference calls ms_ssim_np.MultiScaleSSIM with
    batch = 1, val.py:81-93).  uint8 CUDA tensors -> float64 tensor (N,)."""
    assert img1.dim() == img2.dim() == 4, 'Expected {}'.format(data_format)          # ms_ssim_np.py:26
    assert img1.dtype == torch.uint8, 'Expected uint8 intput'                         # :27
    assert img2.dtype == torch.uint8, 'Expected uint8 intput'                         # :28
    if tuple(img1.shape) != tuple(img2.shape):
        raise RuntimeError('Input images must have the same shape (%s vs. %s).' % (tuple(img1.shape), tuple(img2.shape)))
    if data_format == 'NHWC':
        img1, img2 = img1.permute(0, 3, 1, 2), img2.permute(0, 3, 1, 2)
    assert img1.shape[1] == 3, 'Expected 3-channel images, got {}'.format(tuple(img1.shape))   # :35
    _lib.require_device()
    a, b = img1.contiguous(), img2.contiguous()
    N, _, H, W = a.shape
    L = _lib.lib()
    ws = torch.empty(L.ic_msssim_workspace_bytes(N, H, W, 1), dtype=torch.uint8, device=a.device)
    out = torch.empty(N, dtype=torch.float64, device=a.device)
    _lib.check(L.ic_msssim_np_fwd(_lib.ptr(a), _lib.ptr(b), N, H, W, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                  _lib.stream_ptr()))
    return out


def tf_msssim_np(img1, img2, data_format='NHWC'):
    """code/ms_ssim_np.py:25-40: float32 scalar for a batch of ONE image."""
    v = MultiScaleSSIM_batch(img1, img2, data_format)
    if v.numel() != 1:
        raise ValueError('tf_msssim_np is per image (val.py feeds batch = 1); use MultiScaleSSIM_batch')
    return v[0].float()
