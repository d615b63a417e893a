"""Host mirror of the reference's ``autoencoder`` module (code/autoencoder.py).

Same names / argument order / returned field names as the reference objects that
code/val.py:74-86 and code/train.py:86-102 call, but eager: tensors are torch
CUDA tensors used as containers, all arithmetic runs in libimgcomp_b200.so.
"""
from collections import namedtuple

import numpy as np
import torch

from . import _lib, quantizer

# returned by _Network.encode (code/autoencoder.py:15)
EncoderOutput = namedtuple('EncoderOutput', ['qbar', 'qhard', 'symbols', 'z', 'heatmap'])

SCOPE_AE = 'autoencoder'
SCOPE_AE_ENC = SCOPE_AE + '/encoder'
SCOPE_AE_DEC = SCOPE_AE + '/decoder'


def get_network_cls(config):
    """code/autoencoder.py:26-29"""
    return {'CVPR': _CVPR}[config.arch]


class _Network(object):
    """code/autoencoder.py:32-204.  ``weights``: dict TF-variable-name -> ndarray
    (see weights.py); the reference creates variables inside encode()/decode()
    and restores them from a checkpoint, here they are handed over explicitly
    (constructor or load_weights).  ``mode``: 'exact' (default; tcgen05 fp16x3, float32-class), 'fp32'
    (float32 FFMA kernels, strict parity, ~10x slower) or 'fast' (single fp16 pass) -- DESIGN.md 4.2."""

    def __init__(self, config, quantize=True, weights=None, mode='exact'):
        if not quantize:
            raise NotImplementedError('quantize=False is not on the hot path')
        self.config = config
        self.quantize = quantize
        self.mode = _lib.MODES[mode]
        self.num_chan_bn_including_heatmap = config.num_chan_bn + 1
        self._centers = None       # set in encode(); access with get_centers_variable (autoencoder.py:43)
        self._handle = None
        self._ws = {}
        self.extra = {}            # outputs of the last encode() beyond EncoderOutput (qsoft, symbols_u8)
        self._cfg = _lib.AeConfig(config.num_chan_bn, config.arch_param_B, config.num_centers,
                                  1 if config.heatmap else 0, {'OFF': 0, 'FIXED': 1}[config.normalization])
        if weights is not None:
            self.load_weights(weights)

    # -- weights -----------------------------------------------------------
    def variable_names(self):
        L = _lib.lib()
        n = L.ic_ae_num_tensors(self._cfg)
        _lib.check(min(n, 0))
        return [L.ic_ae_tensor_name(self._cfg, i).decode() for i in range(n)]

    def load_weights(self, weights):
        _lib.require_device()
        L = _lib.lib()
        names = self.variable_names()
        arrays = []
        for i, name in enumerate(names):
            a = np.ascontiguousarray(np.asarray(weights[name], dtype=np.float32))
            if a.size != L.ic_ae_tensor_numel(self._cfg, i):
                raise ValueError('%s has %d elements, expected %d' % (name, a.size, L.ic_ae_tensor_numel(self._cfg, i)))
            arrays.append(a)
        h = _lib.c_void_p()
        _lib.check(L.ic_ae_create(self._cfg, _lib.host_tensor_array(arrays), len(arrays), h))
        if self._handle is not None:
            L.ic_ae_destroy(self._handle)
        self._handle = h
        self._centers_value = torch.from_numpy(np.asarray(weights[SCOPE_AE_ENC + '/centers'], np.float32)).cuda()

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().ic_ae_destroy(self._handle)
        except Exception:
            pass

    def _need_handle(self):
        if self._handle is None:
            raise RuntimeError('no weights loaded: pass weights= to the constructor or call load_weights()')

    def _workspace(self, key, nbytes):
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device='cuda')
            if len(self._ws) > 4:
                self._ws.clear()
            self._ws[key] = ws
        return ws

    # -- reference API -----------------------------------------------------
    @staticmethod
    def get_subsampling_factor():
        raise NotImplementedError()

    def encode(self, x, is_training):
        """-> EncoderOutput(qbar, qhard, symbols, z, heatmap) (code/autoencoder.py:50-58).
        x: CUDA tensor N x 3 x H x W, float32 in [0,255] (or uint8: tf.to_float of
        val.py:83 is fused)."""
        if is_training:
            raise NotImplementedError('is_training=True (batch-statistics BN + backward) is not built yet')
        if x.dtype not in (torch.float32, torch.uint8):
            raise AssertionError('Expected float32 for x, got {}'.format(x.dtype))     # autoencoder.py:51
        if x.dim() != 4 or x.shape[1] != 3 or not x.is_cuda:
            raise ValueError('expected a CUDA tensor N x 3 x H x W, got {}'.format(tuple(x.shape)))
        self._need_handle()
        self._centers = self._centers_value
        return self._encode(x.contiguous(), is_training)

    def decode(self, q, is_training):
        """-> x_out N x 3 x 8h x 8w float32 clipped to [0,255] (code/autoencoder.py:60-63)."""
        if is_training:
            raise NotImplementedError('is_training=True is not built yet')
        self._need_handle()
        return self._decode(q.contiguous().float(), is_training)

    def centers_tensor(self):
        """the loaded centers (L,) float32 CUDA; unlike get_centers_variable it needs no encode() first --
        a decoder process never encodes (codec.py)."""
        self._need_handle()
        return self._centers_value

    def get_centers_variable(self):
        if self._centers is None:
            raise ValueError('Call -encode(...) before trying to access centers')      # autoencoder.py:66-67
        return self._centers


class _CVPR(_Network):
    """code/autoencoder.py:213-268"""

    @staticmethod
    def get_subsampling_factor():
        return 8

    def _encode(self, x, is_training):
        L = _lib.lib()
        N, _, H, W = x.shape
        f = self.get_subsampling_factor()
        if H % f or W % f:
            raise ValueError('H and W must be multiples of {} (val.py:157 pads images), got {}x{}'.format(f, H, W))
        C = self.config.num_chan_bn
        shape = (N, C, H // f, W // f)
        dev = x.device
        z, hm, qbar, qhard, qsoft = (torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(5))
        sym = torch.empty(shape, dtype=torch.int64, device=dev)
        sym8 = torch.empty(shape, dtype=torch.uint8, device=dev)
        nbytes = L.ic_encode_workspace_bytes(self._handle, N, H, W, self.mode)
        ws = self._workspace(('enc', N, H, W), nbytes)
        _lib.check(L.ic_encode_fwd(self._handle, _lib.ptr(x), int(x.dtype == torch.uint8), N, H, W,
                                   _lib.ptr(z), _lib.ptr(hm) if self.config.heatmap else None, _lib.ptr(qbar),
                                   _lib.ptr(qhard), _lib.ptr(sym), _lib.ptr(sym8), _lib.ptr(qsoft),
                                   _lib.ptr(ws), ws.numel(), self.mode, _lib.stream_ptr()))
        self.extra = {'qsoft': qsoft, 'symbols_u8': sym8}
        return EncoderOutput(qbar, qhard, sym, z, hm if self.config.heatmap else None)

    def _decode(self, q, is_training):
        L = _lib.lib()
        N, C, h, w = q.shape
        if C != self.config.num_chan_bn:
            raise ValueError('expected {} channels, got {}'.format(self.config.num_chan_bn, C))
        x_out = torch.empty((N, 3, 8 * h, 8 * w), dtype=torch.float32, device=q.device)
        x_out_u8 = torch.empty((N, 3, 8 * h, 8 * w), dtype=torch.uint8, device=q.device)
        nbytes = L.ic_decode_workspace_bytes(self._handle, N, h, w, self.mode)
        ws = self._workspace(('dec', N, h, w), nbytes)
        _lib.check(L.ic_decode_fwd(self._handle, _lib.ptr(q), N, h, w, _lib.ptr(x_out), _lib.ptr(x_out_u8),
                                   _lib.ptr(ws), ws.numel(), self.mode, _lib.stream_ptr()))
        self.extra['x_out_u8'] = x_out_u8      # tf.cast(x_out, tf.uint8) of val.py:91, fused
        return x_out
