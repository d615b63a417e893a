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

# A trainable variable as the reference's callers see one (tf.Variable: .name / value), code/autoencoder.py:71-79
Variable = namedtuple('Variable', ['name', 'value'])

# The reference keeps variables and regularisation terms in TF's global graph collections, which is why
# encoder_variables() / *_regularization_loss() are static there (code/autoencoder.py:71-88).  Same here: the variables
# of the last network whose weights were loaded, by TF name, and the config that carries the regularisation factors.
_GRAPH = {'variables': {}, 'config': None}


def _trainable(name):
    return not (name.endswith('/moving_mean') or name.endswith('/moving_variance'))


def _l2_loss(w):
    """tf.nn.l2_loss: sum(w ** 2) / 2"""
    return float(np.sum(np.asarray(w, np.float64) ** 2) / 2)


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
        self._centers_host = np.ascontiguousarray(np.asarray(weights[SCOPE_AE_ENC + '/centers'], np.float32))
        self._centers_value = torch.from_numpy(self._centers_host).cuda()
        self._weights = {n: a for n, a in zip(names, arrays)}
        self._train = None
        _GRAPH['variables'] = self._weights
        _GRAPH['config'] = self.config

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
        if x.dtype not in (torch.float32, torch.uint8):
            raise AssertionError('Expected float32 for x, got {}'.format(x.dtype))     # autoencoder.py:51
        if x.dim() != 4 or x.shape[1] != 3 or not x.is_cuda:
            raise ValueError('expected a CUDA tensor N x 3 x H x W, got {}'.format(tuple(x.shape)))
        self._need_handle()
        self._centers = self._centers_value
        if is_training:
            return self._encode_training(x.contiguous())
        return self._encode(x.contiguous(), is_training)

    def decode(self, q, is_training):
        """-> x_out N x 3 x 8h x 8w float32 clipped to [0,255] (code/autoencoder.py:60-63)."""
        self._need_handle()
        if is_training:
            return self._training_net().decode_forward(q.contiguous().float(), is_training=True)
        return self._decode(q.contiguous().float(), is_training)

    # -- is_training=True: batch-statistics batch norm (code/autoencoder.py:115-125), the forward of code/train.py:101-102
    def _training_net(self):
        """The training-mode forward lives in trainer.Trainer (batch-norm batch statistics, float32 NHWC primitives /
        tcgen05 3x3 convs); built lazily over this network's variables.  Forward only: gradients and the optimiser are
        trainer.Trainer.step (train.get_train_op)."""
        if self._train is None:
            from . import config as config_mod, trainer, weights as weights_mod
            pc_cfg = config_mod.pc_config('cvpr/res_shallow')
            W = weights_mod.synthetic_weights(self.config.num_chan_bn, self.config.num_centers, pc_cfg.arch_param__k,
                                              self.config.arch_param_B)      # context-model slots of the Trainer: unused here
            W.update(self._weights)
            self._train = trainer.Trainer(self.config, pc_cfg, W, mode='fp32' if self.mode == _lib.IC_MODE_FP32 else 'exact')
        return self._train

    def _encode_training(self, x):
        enc = self._training_net().encode_forward(x, is_training=True)
        self.extra = {'qsoft': enc['qsoft']}
        return EncoderOutput(enc['qbar'], enc['qhard'], enc['symbols'], enc['z'], enc['heatmap'])

    def centers_tensor(self):
        """the loaded centers (L,) float32 CUDA; unlike get_centers_variable it needs no encode() first --
        a decoder process never encodes (codec.py)."""
        self._need_handle()
        return self._centers_value

    def centers_host(self):
        """the loaded centers as a host float32 array (the codec hands them to calls that must not synchronise)"""
        self._need_handle()
        return self._centers_host

    def get_centers_variable(self):
        if self._centers is None:
            raise ValueError('Call -encode(...) before trying to access centers')      # autoencoder.py:66-67
        return self._centers

    @staticmethod
    def _get_trainable_vars_assert_non_empty(scope):
        """code/autoencoder.py:81-88"""
        v = [Variable(n, a) for n, a in _GRAPH['variables'].items() if n.startswith(scope + '/') and _trainable(n)]
        assert len(v) > 0, 'No trainable vars in scope {}. All: {}'.format(scope, sorted(_GRAPH['variables']))
        return v

    @staticmethod
    def encoder_variables():
        """ Includes center variable (code/autoencoder.py:71-74) """
        return _Network._get_trainable_vars_assert_non_empty(scope=SCOPE_AE_ENC)

    @staticmethod
    def decoder_variables():
        """code/autoencoder.py:76-78"""
        return _Network._get_trainable_vars_assert_non_empty(scope=SCOPE_AE_DEC)

    @staticmethod
    def _regularization_loss(scope):
        """tf.losses.get_regularization_loss(scope): slim.l2_regularizer(regularization_factor) on every conv weight of
        the scope (code/autoencoder.py:98-102) + regularization_factor_centers * l2_loss(centers) (code/quantizer.py:18-24)"""
        cfg = _GRAPH['config']
        assert cfg is not None, 'no variables: load weights first'
        total = 0.0
        for n, a in _GRAPH['variables'].items():
            if not n.startswith(scope + '/'):
                continue
            if n.endswith('/weights'):
                total += cfg.regularization_factor * _l2_loss(a)
            elif n.endswith('/centers') and cfg.regularization_factor_centers != 0:
                total += cfg.regularization_factor_centers * _l2_loss(a)
        return total

    @staticmethod
    def encoder_regularization_loss():
        """ includes centers regularization (code/autoencoder.py:80-83) """
        return _Network._regularization_loss(SCOPE_AE_ENC)

    @staticmethod
    def decoder_regularization_loss():
        """code/autoencoder.py:85-87"""
        return _Network._regularization_loss(SCOPE_AE_DEC)


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
