"""Host mirror of code/probclass.py: the masked causal 3-D-conv context model
and the helpers the --real_bpp path uses."""
import itertools

import numpy as np
import torch

from . import _lib


def get_network_cls(pc_config):
    """Returns a class that is a subclass of _Network3D (code/probclass.py:11-15)."""
    return {'res_shallow': _ResShallow}[pc_config.arch]


def context_shape_from_context_size(context_size):
    """:return context shape as DHW (code/probclass.py:18-20)"""
    return context_size // 2 + 1, context_size, context_size


def context_size_from_context_shape(context_shape):
    return context_shape[-1]


class _Network3D(object):
    _PROBCLASS_SCOPE = 'probclass3d'

    def __init__(self, pc_config, num_centers, weights=None):
        """code/probclass.py:30-41.  ``weights``: dict TF-variable-name -> ndarray."""
        self.config = pc_config
        self.reuse = False
        self.L = num_centers
        self._handle = None
        self._ws = None
        if getattr(pc_config, 'learn_pad_var', False):
            raise NotImplementedError('learn_pad_var=True is unused by the published configs (pc_configs/base:24)')
        self._cfg = _lib.PcConfig(pc_config.kernel_size, pc_config.arch_param__k, num_centers)
        if weights is not None:
            self.load_weights(weights)

    @classmethod
    def get_num_layers(cls):
        raise NotImplementedError()

    @classmethod
    def get_context_size(cls, config):
        """width / height of the receptive field (code/probclass.py:47-52)"""
        return cls.get_num_layers() * (config.kernel_size - 1) + 1

    @classmethod
    def get_context_shape(cls, config):
        """Shape as DHW (code/probclass.py:54-57)"""
        return context_shape_from_context_size(cls.get_context_size(config))

    def auto_pad_value(self, ae):
        """code/probclass.py:59-61"""
        return 0 if not self.config.use_centers_for_padding else ae.get_centers_variable()[0]

    _pad_cache = (None, None)

    def _pad_float(self, pad_value):
        """The C ABI takes the pad value as a host float; a device scalar (centers[0]) is read back once per
        version of its storage, not once per call (a read-back per call would drain the stream every step)."""
        if not torch.is_tensor(pad_value):
            return float(pad_value)
        key = (pad_value.data_ptr(), pad_value._version, pad_value.device)
        if self._pad_cache[0] != key:
            self._pad_cache = (key, float(pad_value))
        return self._pad_cache[1]

    # -- weights -----------------------------------------------------------
    def variable_names(self):
        L = _lib.lib()
        n = L.ic_pc_num_tensors(self._cfg)
        _lib.check(min(n, 0))
        return [L.ic_pc_tensor_name(self._cfg, i).decode() for i in range(n)]

    def load_weights(self, weights):
        _lib.require_device()
        L = _lib.lib()
        arrays = []
        for i, name in enumerate(self.variable_names()):
            a = np.ascontiguousarray(np.asarray(weights[name], dtype=np.float32))
            if a.size != L.ic_pc_tensor_numel(self._cfg, i):
                raise ValueError('%s has %d elements, expected %d' % (name, a.size, L.ic_pc_tensor_numel(self._cfg, i)))
            arrays.append(a)
        h = _lib.c_void_p()
        _lib.check(L.ic_pc_create(self._cfg, _lib.host_tensor_array(arrays), len(arrays), h))
        if self._handle is not None:
            L.ic_pc_destroy(self._handle)
        self._handle = h
        self._weights = dict(zip(self.variable_names(), arrays))

    def variables(self):
        """tf.trainable_variables('probclass3d') (code/probclass.py:109-114) as (name, value) pairs"""
        from .autoencoder import Variable
        trainable_vars = [Variable(n, a) for n, a in getattr(self, '_weights', {}).items()]
        assert len(trainable_vars) > 0, 'No trainable variables found in scope {}.'.format(self._PROBCLASS_SCOPE)
        return trainable_vars

    def regularization_loss(self):
        """code/probclass.py:116-120: None unless pc_config.regularization_factor is set; then
        regularization_factor * sum l2_loss(conv3d weights) (the regulariser of :249-251)"""
        if self.config.regularization_factor is None:
            return None
        return self.config.regularization_factor * sum(
            float(np.sum(np.asarray(a, np.float64) ** 2) / 2) for n, a in self._weights.items() if n.endswith('/weights'))

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().ic_pc_destroy(self._handle)
        except Exception:
            pass

    def _need_handle(self):
        if self._handle is None:
            raise RuntimeError('no weights loaded: pass weights= to the constructor or call load_weights()')

    def _workspace(self, N, D, H, W):
        nbytes = _lib.lib().ic_pc_workspace_bytes(self._handle, N, D, H, W)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device='cuda')
        return self._ws

    # -- reference API -----------------------------------------------------
    def bitcost(self, q, target_symbols, is_training, pad_value=0):
        """Pads q, runs the context model, cross entropy against target_symbols.
        q NCHW float32, target_symbols NCHW int64 -> bitcost per symbol NCHW
        (code/probclass.py:63-106)."""
        # is_training only reaches _logits, whose conv3d layers have no batch norm / dropout (code/probclass.py:214-261):
        # the training-mode forward IS this forward.  Gradients: trainer.Trainer.step (train.get_train_op).
        assert q.dim() == 4                                                      # tf_helpers.assert_ndims(q, 4), :71
        self._need_handle()
        self.reuse = True
        q = q.contiguous().float()
        sym = target_symbols.contiguous().to(torch.int64)
        assert tuple(sym.shape) == tuple(q.shape)                                # :98
        N, C, h, w = q.shape
        bits = torch.empty_like(q)
        sums = torch.empty(N, dtype=torch.float64, device=q.device)
        ws = self._workspace(N, C, h, w)
        _lib.check(_lib.lib().ic_pc_bitcost_fwd(self._handle, _lib.ptr(q), _lib.ptr(sym), self._pad_float(pad_value), N, C, h, w,
                                                _lib.ptr(bits), _lib.ptr(sums), _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr()))
        self.last_bits_per_image = sums
        return bits

    def logits(self, q, is_training=False):
        """For accessing logits on an UN-padded NDHW1 volume (code/probclass.py:130-135)."""
        assert self.reuse, 'Make sure to call bitcost(...) before calling logits(...)'     # :132
        self._need_handle()
        if q.dim() == 5:
            assert q.shape[-1] == 1
            q = q[..., 0]
        q = q.contiguous().float()
        N, D, H, W = q.shape
        out = torch.empty((N, D - 4, H - 8, W - 8, self.L), dtype=torch.float32, device=q.device)
        ws = self._workspace(N, D, H, W)
        _lib.check(_lib.lib().ic_pc_logits_fwd(self._handle, _lib.ptr(q), N, D, H, W, _lib.ptr(out), _lib.ptr(ws),
                                               ws.numel(), _lib.stream_ptr()))
        return out

    def freqs(self, symbols, centers, codec=False):
        """ONE batched pass replacing the reference's per-symbol PredictionNetwork loop
        (code/probclass.py:441-476 driven by code/bit_counter.py:103-134).
        symbols NCHW int64 -> (freqs N,C,h,w,L int64 ; theoretical bits per image (N,) float64).
        codec=True: the tables of a real bitstream, in the arithmetic decode_streams() reproduces."""
        self._need_handle()
        sym = symbols.contiguous().to(torch.int64)
        N, C, h, w = sym.shape
        out = torch.empty((N, C, h, w, self.L), dtype=torch.int64, device=sym.device)
        sums = torch.empty(N, dtype=torch.float64, device=sym.device)
        ws = self._workspace(N, C, h, w)
        centers = centers.contiguous().float()
        fn = _lib.lib().ic_pc_codec_freqs_fwd if codec else _lib.lib().ic_pc_freqs_fwd
        _lib.check(fn(self._handle, _lib.ptr(sym), _lib.ptr(centers), N, C, h, w, _lib.ptr(out),
                      _lib.ptr(sums), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return out, sums

    def codec_freqs_u32(self, symbols, centers_host, out=None):
        """The tables of a real bitstream (freqs(codec=True)) as uint32, enqueued without any host synchronisation:
        centers_host is a HOST array (L floats).  symbols NCHW int64 CUDA -> N,C,h,w,L uint32 CUDA (`out` if given)."""
        self._need_handle()
        sym = symbols.contiguous().to(torch.int64)
        N, C, h, w = sym.shape
        if out is None:
            out = torch.empty((N, C, h, w, self.L), dtype=torch.int32, device=sym.device)
        assert out.is_cuda and out.numel() == sym.numel() * self.L and out.element_size() == 4
        ch = np.ascontiguousarray(np.asarray(centers_host, np.float32))
        assert ch.size == self.L
        ws = self._workspace(N, C, h, w)
        _lib.check(_lib.lib().ic_pc_codec_freqs_u32_fwd(self._handle, _lib.ptr(sym), ch.ctypes.data, N, C, h, w, _lib.ptr(out),
                                                        None, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return out

    def decode_streams(self, streams, first_syms, shape, centers, force_symbols=None, return_freqs=False):
        """The decoder half of --real_bpp (code/bit_counter.py:137-163), for N images at once:
        streams = list of N byte strings written by ArithmeticEncoder over freqs(codec=True) tables,
        first_syms = the N side-information symbols, shape = (C, h, w).
        -> symbols N,C,h,w uint8 (CUDA).  One CTA per image: cached-activation context model +
        range decoder on the device.  force_symbols / return_freqs are debug hooks."""
        self._need_handle()
        L = _lib.lib()
        N = len(streams)
        C, h, w = (int(v) for v in shape)
        offs = np.zeros(N + 1, np.int64)
        offs[1:] = np.cumsum([len(b) for b in streams])
        blob = np.frombuffer(b''.join(bytes(b) for b in streams) + b'\0' * 16, dtype=np.uint8)
        d_stream = torch.from_numpy(blob.copy()).cuda()
        d_offs = torch.from_numpy(offs).cuda()
        d_first = torch.tensor([int(v) for v in first_syms], dtype=torch.int32, device='cuda')
        out = torch.zeros((N, C, h, w), dtype=torch.uint8, device='cuda')
        force = None
        if force_symbols is not None:
            force = force_symbols.to('cuda', torch.uint8).contiguous()
            assert tuple(force.shape) == (N, C, h, w)
        seen = torch.zeros((N, C, h, w, self.L), dtype=torch.int64, device='cuda') if return_freqs else None
        nbytes = L.ic_pc_decode_workspace_bytes(self._handle, N, C, h, w)
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device='cuda')
        centers = centers.contiguous().float()
        _lib.check(L.ic_pc_decode_fwd(self._handle, _lib.ptr(d_stream), _lib.ptr(d_offs), _lib.ptr(d_first), _lib.ptr(centers),
                                      N, C, h, w, _lib.ptr(out), _lib.ptr(force) if force is not None else None,
                                      _lib.ptr(seen) if seen is not None else None, _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr()))
        return (out, seen) if return_freqs else out


class _ResShallow(_Network3D):
    """code/probclass.py:199-221"""
    _NUM_RESIDUAL = 1

    @classmethod
    def get_num_layers(cls):
        return 2 + _ResShallow._NUM_RESIDUAL * 2


# Helpers ----------------------------------------------------------------------

def pad_for_probclass3d(x, context_size, pad_value=0, learn_pad_var=False):
    """code/probclass.py:268-292 for numpy CHW / NCHW symbol volumes (bit_counter use)."""
    assert not learn_pad_var
    pad = context_size // 2
    assert pad >= 1
    pads = [(0, 0)] * (x.ndim - 3) + [(pad, 0), (pad, pad), (pad, pad)]
    return np.pad(x, pads, mode='constant', constant_values=pad_value)


def undo_pad_for_probclass3d(x, context_size):
    """code/probclass.py:341-351"""
    pad = context_size // 2
    return x[..., pad:, pad:-pad, pad:-pad]


def _iter_block_idices(syms_shape, block_sizes):
    """raster order C -> H -> W (code/probclass.py:382-387)"""
    C, H, W = syms_shape
    bC, bH, bW = block_sizes
    for c, h, w in itertools.product(range(C - bC + 1), range(H - bH + 1), range(W - bW + 1)):
        yield slice(c, c + bC), slice(h, h + bH), slice(w, w + bW)


def iter_over_blocks(syms, block_sizes):
    """code/probclass.py:367-375"""
    for c_slice, h_slice, w_slice in _iter_block_idices(syms.shape, block_sizes):
        yield syms[c_slice, h_slice, w_slice]


def num_blocks(syms_shape, block_sizes):
    C, H, W = syms_shape
    bC, bH, bW = block_sizes
    return (C - bC + 1) * (H - bH + 1) * (W - bW + 1)


class ProbclassNetworkTesting(object):
    """Bit cost of a symbol volume, fully convolutionally (code/probclass.py:393-421)."""

    def __init__(self, pc, ae, sess=None):
        self.pc, self.ae = pc, ae

    def get_total_bit_cost(self, symbols):
        """:param symbols: CHW or NCHW (numpy or tensor) :return: total bits"""
        s = torch.as_tensor(np.asarray(symbols) if not torch.is_tensor(symbols) else symbols)
        if s.dim() == 3:
            s = s[None]
        assert s.dim() == 4
        s = s.to('cuda', torch.int64)
        centers = self.ae.get_centers_variable()
        q = centers[s]                                      # tf.gather(centers, symbols), :407
        bc = self.pc.bitcost(q, s, is_training=False, pad_value=self.pc.auto_pad_value(self.ae))
        return float(bc.sum(dtype=torch.float64))


class PredictionNetwork(object):
    """Prediction given (slices of) the symbol volume (code/probclass.py:425-482).
    New: get_all_freqs evaluates every context of a volume in one pass."""

    def __init__(self, pc, config, centers, sess=None, freqs_resolution=1e9):
        assert freqs_resolution == 1e9, 'the kernel bakes in the reference default 1e9'
        self.pc = pc
        self.pc_class = pc.__class__
        self.config = config
        self.centers = centers
        self.input_ctx_shape = self.pc_class.get_context_shape(config)

    def pad_symbols_volume(self, symbols):
        assert symbols.ndim == 3
        return pad_for_probclass3d(symbols, self.pc_class.get_context_size(self.config))

    def undo_pad_symbols_volume(self, symbols):
        assert symbols.ndim == 3
        return undo_pad_for_probclass3d(symbols, self.pc_class.get_context_size(self.config))

    def _ctx_logits(self, input_ctx):
        ctx = np.asarray(input_ctx)
        assert ctx.shape == tuple(self.input_ctx_shape), '{} != {}'.format(ctx.shape, self.input_ctx_shape)
        s = torch.from_numpy(ctx.astype(np.int64)).cuda()
        q = self.centers[s][None]                           # tf.gather(centers, ctx), 1DHW
        self.pc.reuse = True
        return self.pc.logits(q)[0, 0, 0, 0]

    def get_pr(self, input_ctx):
        """input_ctx: symbols CHW (5,9,9) -> softmax probabilities (L,) (code/probclass.py:453-459)"""
        return torch.softmax(self._ctx_logits(input_ctx), -1).cpu().numpy()

    def get_freqs(self, input_ctx):
        """input_ctx: symbols CHW (5,9,9) -> int64 freqs (L,) (code/probclass.py:461-476).
        Same kernel arithmetic as get_all_freqs: bit-identical to the batched table."""
        ctx = np.asarray(input_ctx)
        assert ctx.shape == tuple(self.input_ctx_shape), '{} != {}'.format(ctx.shape, self.input_ctx_shape)
        s = torch.from_numpy(np.ascontiguousarray(ctx.astype(np.int64))).cuda()
        pc = self.pc
        pc._need_handle()
        out = torch.empty(pc.L, dtype=torch.int64, device='cuda')
        ws = pc._workspace(1, *ctx.shape)
        centers = self.centers.contiguous().float()
        _lib.check(_lib.lib().ic_pc_context_freqs_fwd(pc._handle, _lib.ptr(s), _lib.ptr(centers), 1, ctx.shape[0],
                                                      ctx.shape[1], ctx.shape[2], _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                                      _lib.stream_ptr()))
        f = out.cpu().numpy()
        assert np.all(f > 0), 'We do not want zero frequencies!: {}'.format(f)
        return f

    def get_all_freqs(self, symbols, codec=False):
        """symbols CHW (numpy or tensor) -> int64 (C,h,w,L) numpy, every position's table in
        the coder's raster order, plus the theoretical bit cost.  codec=True: the tables
        decode_symbols() re-derives bit for bit (use these to write a stream)."""
        s = torch.as_tensor(np.asarray(symbols) if not torch.is_tensor(symbols) else symbols)
        assert s.dim() == 3
        f, bits = self.pc.freqs(s[None].to('cuda', torch.int64), self.centers, codec=codec)
        return f[0].cpu().numpy(), float(bits[0])

    def decode_symbols(self, stream, first_sym, shape):
        """bitstream + side information -> symbols CHW (numpy int64): code/bit_counter.py:137-163
        without the per-symbol sess.run."""
        out = self.pc.decode_streams([stream], [first_sym], shape, self.centers)
        return out[0].cpu().numpy().astype(np.int64)
