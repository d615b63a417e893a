"""Host mirror of code/quantizer.py."""
import torch

from . import _lib


def quantize(x, centers, sigma):
    """:return qsoft, qhard, symbols  (code/quantizer.py:37-40)
    x float32 rank-4 (NCHW), centers float32 (L,)."""
    assert x.dtype == torch.float32, 'x should be float32'                       # quantizer.py:48
    assert centers.dtype == torch.float32, 'centers should be float32'           # :49
    assert x.dim() == 4, 'x should be NCHW or NHWC, got {}'.format(tuple(x.shape))   # :50
    assert centers.dim() == 1, 'centers should be (L,), got {}'.format(tuple(centers.shape))  # :51
    _lib.require_device()
    x = x.contiguous()
    centers = centers.contiguous().to(x.device)
    qsoft, qhard = torch.empty_like(x), torch.empty_like(x)
    symbols = torch.empty(x.shape, dtype=torch.int64, device=x.device)
    _lib.check(_lib.lib().ic_quantize_fwd(_lib.ptr(x), _lib.ptr(centers), centers.numel(), float(sigma), x.numel(),
                                          _lib.ptr(qsoft), _lib.ptr(qhard), _lib.ptr(symbols), _lib.stream_ptr()))
    return qsoft, qhard, symbols
