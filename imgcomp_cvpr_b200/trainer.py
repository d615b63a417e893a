"""The training step of code/train.py (cfg 3: forward + MS-SSIM / rate loss + backward + two Adam groups) on the
float32 training primitives of libimgcomp_b200.so (include/imgcomp_b200.h, ic_nn_* / ic_msssim_tf_bwd).

What the reference builds as a TF graph and differentiates with tf.gradients (code/train.py:86-132,303-349):

    enc   = ae.encode(x, is_training=True)                                   autoencoder.py:218-244
    x_out = ae.decode(enc.qbar, is_training=True)                            :246-268
    bc    = pc.bitcost(stop_gradient(enc.qbar), enc.symbols, True, centers[0])   train.py:103-105
    loss  = K_ms_ssim (1 - MS-SSIM(x, x_out)) + beta max(0.5 (mean(bc heatmap) + mean(bc)) - H_target, 0) + l2 terms
    Adam_AE (lr_ae, staircase decay) on the autoencoder variables, Adam_PC (lr_pc) on probclass3d/*

is executed here eagerly: every op below is one or two kernels of the library, a small tape records the backward
closures in the order of the reference graph, parameters / gradients / Adam moments live in flat per-group buffers
(one Adam launch per group).  torch tensors are containers only.  No CPU fallback.
"""
import math
import os

import numpy as np
import torch

from . import _lib, nn, weights as weights_mod

ARCH_N = weights_mod.ARCH_N


def _ceil4(n):
    return (n + 3) // 4 * 4


def _pc_masks():
    """create_first_mask / create_other_mask of code/probclass.py:150-176 for kernel_size 3 -> two (2,3,3) 0/1 arrays:
    filter depth 0 (previous channel): all 9 taps; depth 1 (current channel): row 0 and (1,0) [first] (+ (1,1) [other])."""
    first = np.zeros((2, 3, 3), np.float32)
    first[0] = 1
    first[1, 0, :] = 1
    first[1, 1, 0] = 1
    other = first.copy()
    other[1, 1, 1] = 1
    return first, other


class _Flat(object):
    """Named float32 tensors packed into one flat buffer (+ gradient and Adam moment buffers of the same layout)."""

    def __init__(self, with_opt=True):
        self.entries = []          # (key, device-shape ndarray)
        self.with_opt = with_opt

    def add(self, key, arr):
        self.entries.append((key, np.ascontiguousarray(arr, np.float32)))

    def finish(self, device):
        off, self.slots = 0, {}
        for key, arr in self.entries:
            self.slots[key] = (off, arr.shape)
            off += (arr.size + 63) // 64 * 64           # 256-byte aligned views (float4 loads in the kernels)
        host = np.zeros(max(off, 64), np.float32)
        for key, arr in self.entries:
            o, _ = self.slots[key]
            host[o:o + arr.size] = arr.ravel()
        self.w = torch.from_numpy(host).to(device)
        if self.with_opt:
            self.g = torch.zeros_like(self.w)
            self.m = torch.zeros_like(self.w)
            self.v = torch.zeros_like(self.w)
        del self.entries

    def view(self, buf, key):
        o, shape = self.slots[key]
        return buf[o:o + int(np.prod(shape))].view(shape)


class _Tape(object):
    def __init__(self):
        self.fns, self.grads = [], {}

    def add(self, fn):
        self.fns.append(fn)

    def acc(self, t, g, owned=False):
        """accumulate g into the gradient of t; tensors that are not `owned` are never modified in place"""
        k = id(t)
        if k not in self.grads:
            self.grads[k] = (t, g, owned)
            return
        _, old, old_owned = self.grads[k]
        if old_owned:
            nn.axpby(1.0, old, 1.0, g, out=old)
        elif owned:
            nn.axpby(1.0, g, 1.0, old, out=g)
            self.grads[k] = (t, g, True)
        else:
            self.grads[k] = (t, nn.add(old, g), True)

    def pop(self, t):
        e = self.grads.pop(id(t), None)
        return None if e is None else e[1]

    def take(self, t):
        """the float32 gradient accumulated for t so far (removed from the tape), or None"""
        e = self.grads.get(id(t))
        if e is None or not torch.is_tensor(e[1]):
            return None
        del self.grads[id(t)]
        return e[1]

    def backward(self):
        for fn in reversed(self.fns):
            fn()
        self.fns = []


def ae_layer_table(C, B):
    """(scope, k, stride, cin, cout, transposed) of every autoencoder conv (code/autoencoder.py:218-268)"""
    n = ARCH_N
    E, D = 'autoencoder/encoder', 'autoencoder/decoder'
    enc_res, dec_res = weights_mod.conv_scopes(B)
    t = [(E + '/h1', 5, 2, 3, n // 2, False), (E + '/h2', 5, 2, n // 2, n, False)]
    t += [(s, 3, 1, n, n, False) for s in enc_res]
    t += [(E + '/to_bn', 5, 2, n, C + 1, False), (D + '/from_bn', 3, 2, C, n, True)]
    t += [(s, 3, 1, n, n, False) for s in dec_res]
    t += [(D + '/h12', 5, 2, n, n // 2, True), (D + '/h13', 5, 2, n // 2, 3, True)]
    return t


PC_LAYERS = (('probclass3d/logits/conv3d_conv0_mask', 'first'), ('probclass3d/logits/res1/conv3d_conv1_mask', 'other'),
             ('probclass3d/logits/res1/conv3d_conv2_mask', 'other'), ('probclass3d/logits/conv3d_conv2_mask', 'other'))


def learning_rate_at(cfg, global_step, num_itr_per_epoch):
    """training_helpers.create_learning_rate_tensor (code/training_helpers.py:22-35): FIXED, or exponential decay by
    lr_schedule_decay_rate every lr_schedule_decay_interval epochs (staircase unless the config says otherwise)"""
    if cfg.lr_schedule == 'FIXED':
        return cfg.lr_initial
    p = global_step / float(num_itr_per_epoch * cfg.lr_schedule_decay_interval)
    if cfg.lr_schedule_decay_staircase:
        p = math.floor(p)
    return cfg.lr_initial * cfg.lr_schedule_decay_rate ** p


class Trainer(object):
    """One object = the variables of code/train.py's graph + its train_op.

        tr = Trainer(ae_config, pc_config, weights, num_itr_per_epoch)
        out = tr.step(x)            # x N,3,H,W uint8/float32 CUDA; one sess.run([train_op, ...]) of the reference
        W = tr.weights()            # dict TF variable name -> ndarray (loads into autoencoder / probclass unchanged)
    """

    def __init__(self, ae_config, pc_config, weights, num_itr_per_epoch=1000, device='cuda', mode='exact'):
        """mode 'fp32': every kernel float32 FFMA (strict parity mode); 'exact': forward and data gradient of the 64
        3x3 128->128 convs AND their filter gradients on tcgen05 kernels in fp16 hi/lo arithmetic (float32-class,
        DESIGN.md 4.2 / 4.6)."""
        self.device = torch.device(device)
        if self.device.type == 'cuda':
            _lib.require_device()        # device='cpu' builds the parameter buffers only (packing / export logic; no kernels)
        assert mode in ('fp32', 'exact')
        self.mode = mode
        assert ae_config.arch == 'CVPR' and pc_config.arch == 'res_shallow'
        assert ae_config.normalization == 'FIXED'
        assert ae_config.optimizer == 'ADAM' and pc_config.optimizer == 'ADAM', 'only the Adam of the published configs is built'
        assert not pc_config.learn_pad_var
        self.ae_config, self.pc_config = ae_config, pc_config
        self.C, self.B, self.L = ae_config.num_chan_bn, ae_config.arch_param_B, ae_config.num_centers
        self.k = pc_config.arch_param__k
        self.heatmap = bool(ae_config.heatmap)
        self.num_itr_per_epoch = num_itr_per_epoch
        self.global_step = 0
        self.table = ae_layer_table(self.C, self.B)
        self._tr = {s: tr for s, _, _, _, _, tr in self.table}
        self._shape = {s: (k, ci, co) for s, k, _, ci, co, _ in self.table}
        self._stride = {s: st for s, _, st, _, _, _ in self.table}
        # parameter groups: (l2 factor, which optimiser)
        self.groups = {'ae_w': _Flat(), 'ae_bn': _Flat(), 'ae_centers': _Flat(), 'pc_w': _Flat(), 'pc_b': _Flat()}
        self.stats = _Flat(with_opt=False)          # BN moving averages (not optimised)
        self.consts = _Flat(with_opt=False)
        for s, k, _, ci, co, tr in self.table:
            w = np.asarray(weights[s + '/weights'], np.float32)
            if tr:
                assert w.shape == (k, k, co, ci), (s, w.shape)
                w = w.transpose(0, 1, 3, 2)                               # -> [kh][kw][Cin of the op][Cout of the op]
            assert w.shape == (k, k, ci, co), (s, w.shape)
            wp = np.zeros((k, k, _ceil4(ci), _ceil4(co)), np.float32)
            wp[:, :, :ci, :co] = w
            self.groups['ae_w'].add(s + '/weights', wp)
            for name, fill, grp in (('gamma', 0.0, self.groups['ae_bn']), ('beta', 0.0, self.groups['ae_bn']),
                                    ('moving_mean', 0.0, self.stats), ('moving_variance', 1.0, self.stats)):
                v = np.full(_ceil4(co), fill, np.float32)
                v[:co] = np.asarray(weights[s + '/BatchNorm/' + name], np.float32)
                grp.add(s + '/BatchNorm/' + name, v)
        self.groups['ae_centers'].add('autoencoder/encoder/centers', np.asarray(weights['autoencoder/encoder/centers'], np.float32))
        first, other = _pc_masks()
        cin = 1
        self._pc_chans = {}
        for i, (s, mk) in enumerate(PC_LAYERS):
            cout = self.k if i < 3 else self.L
            self._pc_chans[s] = (cin, cout)
            w = np.asarray(weights[s + '/weights'], np.float32)
            assert w.shape == (2, 3, 3, cin, cout), (s, w.shape)
            wp = np.zeros((2, 3, 3, _ceil4(cin), _ceil4(cout)), np.float32)
            wp[..., :cin, :cout] = w
            self.groups['pc_w'].add(s + '/weights', wp)
            b = np.zeros(_ceil4(cout), np.float32)
            b[:cout] = np.asarray(weights[s + '/biases'], np.float32)
            self.groups['pc_b'].add(s + '/biases', b)
            m = np.zeros_like(wp)
            m[..., :cin, :cout] = (first if mk == 'first' else other)[:, :, :, None, None]
            self.consts.add(s + '/mask', m)
            cin = cout
        self.consts.add('ones', np.ones(1024, np.float32))
        self.consts.add('zeros', np.zeros(1024, np.float32))
        for g in list(self.groups.values()) + [self.stats, self.consts]:
            g.finish(self.device)
        self._group_of = {}
        for gname, g in self.groups.items():
            for key in g.slots:
                self._group_of[key] = gname
        # the 3x3 128 -> 128 trunk convs: their power-of-two weight scales are computed by ONE launch per step
        self._trunk = [s for s, k, _, ci, co, _ in self.table if (k, ci, co) == (3, 128, 128)]
        self._trunk_row = {s: i for i, s in enumerate(self._trunk)}
        if self.device.type == 'cuda' and self._trunk:
            g = self.groups['ae_w']
            offs = [g.view(g.w, s + '/weights').data_ptr() - g.w.data_ptr() for s in self._trunk]
            self._trunk_off = torch.tensor([o // 4 for o in offs], dtype=torch.int64, device=self.device)
            self._trunk_scales = torch.zeros(len(self._trunk), 4, dtype=torch.float32, device=self.device)
        self._planes_of = {}
        self._stats_of = None
        self._bn_partial = None
        self._prep = self._prep_buf = None
        self._prep_w = -1
        self.tape = None
        self._adam_t = 0
        self._graph = None
        self._sums = torch.zeros(8, dtype=torch.float64, device=self.device)     # [sum bc, sum bc*hm, (sum w, sum w^2) x 3 groups]
        self._coef = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._val = None
        self._loss_ws = torch.empty(_lib.lib().ic_loss_workspace_bytes(), dtype=torch.uint8, device=self.device)
        self._lr_host = torch.zeros(2, dtype=torch.float32)
        if self.device.type == 'cuda':
            self._lr_host = self._lr_host.pin_memory()
        self._lr_dev = torch.zeros(2, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ parameter access
    def _w(self, key):
        g = self.groups[self._group_of[key]]
        return g.view(g.w, key)

    def _g(self, key):
        g = self.groups[self._group_of[key]]
        return g.view(g.g, key)

    def _const(self, key, n=None):
        v = self.consts.view(self.consts.w, key)
        return v if n is None else v[:n]

    def _export(self, buf_name):
        out = {}
        for s, k, _, ci, co, tr in self.table:
            g = self.groups['ae_w']
            w = g.view(getattr(g, buf_name), s + '/weights')[:, :, :ci, :co].cpu().numpy()
            out[s + '/weights'] = np.ascontiguousarray(w.transpose(0, 1, 3, 2) if tr else w)
            g = self.groups['ae_bn']
            for name in ('gamma', 'beta'):
                out[s + '/BatchNorm/' + name] = g.view(getattr(g, buf_name), s + '/BatchNorm/' + name)[:co].cpu().numpy()
        g = self.groups['ae_centers']
        out['autoencoder/encoder/centers'] = g.view(getattr(g, buf_name), 'autoencoder/encoder/centers').cpu().numpy()
        cin = 1
        for i, (s, _) in enumerate(PC_LAYERS):
            cout = self.k if i < 3 else self.L
            g = self.groups['pc_w']
            out[s + '/weights'] = g.view(getattr(g, buf_name), s + '/weights')[..., :cin, :cout].cpu().numpy()
            g = self.groups['pc_b']
            out[s + '/biases'] = g.view(getattr(g, buf_name), s + '/biases')[:cout].cpu().numpy()
            cin = cout
        return out

    def weights(self):
        """dict TF variable name -> ndarray, in the reference's shapes"""
        out = self._export('w')
        for s, _, _, _, co, _ in self.table:
            for name in ('moving_mean', 'moving_variance'):
                out[s + '/BatchNorm/' + name] = self.stats.view(self.stats.w, s + '/BatchNorm/' + name)[:co].cpu().numpy()
        return out

    def state_dict(self):
        """Everything needed to resume (the role of the reference's Saver, code/saver.py / restore_manager.py, without the
        TF checkpoint format): the variables under their TF names, both Adam moments under '<name>/Adam' and '<name>/Adam_1'
        (tf.train.AdamOptimizer's slot names) and the step counters."""
        out = self.weights()
        for slot, buf in (('Adam', 'm'), ('Adam_1', 'v')):
            for k, v in self._export(buf).items():
                out[k + '/' + slot] = v
        out['global_step'] = np.asarray(self.global_step, np.int64)
        out['adam_step'] = np.asarray(self._adam_t, np.int64)
        return out

    def load_state_dict(self, state):
        """inverse of state_dict(); variables that are missing keep their current value (restore_skip_vars-like)"""
        fresh = Trainer(self.ae_config, self.pc_config, {k: state.get(k, v) for k, v in self.weights().items()},
                        self.num_itr_per_epoch, str(self.device), self.mode)
        for name in self.groups:
            self.groups[name].w.copy_(fresh.groups[name].w)
        self.stats.w.copy_(fresh.stats.w)
        for slot, buf in (('Adam', 'm'), ('Adam_1', 'v')):
            sub = {k[:-len('/' + slot)]: v for k, v in state.items() if k.endswith('/' + slot)}
            if sub:
                tmp = Trainer(self.ae_config, self.pc_config, {k: sub.get(k, np.zeros_like(v)) for k, v in self.weights().items()},
                              self.num_itr_per_epoch, str(self.device), self.mode)
                for name in self.groups:
                    getattr(self.groups[name], buf).copy_(tmp.groups[name].w)
        self.global_step = int(state.get('global_step', self.global_step))
        self._adam_t = int(state.get('adam_step', self._adam_t))
        return self

    def gradients(self):
        """d total_loss / d variable of the last forward_backward(), WITHOUT the l2 terms (those are added inside the
        Adam kernel as l2 * w), same naming / shapes as weights()"""
        return self._export('g')

    # ------------------------------------------------------------------ ops (forward + recorded backward)
    def _tc_plans(self, kind, cin, cout):
        """(forward, data-gradient, filter-gradient plan) of a non-3x3 conv on the tcgen05 kernels, None where no kernel covers the shape
        or outside mode 'exact' (IC_TRAIN_TC_EXTRA=0 keeps these layers on the FFMA kernels: A/B in tests)"""
        if self.mode != 'exact' or os.environ.get('IC_TRAIN_TC_EXTRA', '1') == '0':
            return None, None, None
        return nn.TcPlan.get(kind, False, cin, cout), nn.TcPlan.get(kind, True, cin, cout), nn.TcWgradPlan.get(kind, cin, cout)

    def _fused(self):
        """trunk layers fused (batch norm writes the next conv's planes, the conv accumulates the next batch norm's
        statistics); IC_TRAIN_FUSED=0 keeps the unfused sequence (A/B in tests)"""
        return self.mode == 'exact' and os.environ.get('IC_TRAIN_FUSED', '1') != '0'

    def _begin_pass(self):
        """start of a forward pass: forget the planes of the previous pass, refresh the trunk convs' weight scales"""
        self._planes_of = {}
        self._stats_of = None
        self._prep = None           # packed trunk weights of this pass (filled at the first trunk conv: needs the image width)
        if self._fused() and self._trunk:
            nn.weight_scales(self.groups['ae_w'].w, self._trunk_off, 9 * 128 * 128, self._trunk_scales)

    def _conv(self, x, w, gw, stride=1, transposed=False, valid=False, need_dx=True, mask=None, chans=None, scope=None):
        tc = (self.mode == 'exact' and tuple(w.shape) == (3, 3, 128, 128) and stride == 1 and not transposed and not valid)
        plan_f = plan_d = plan_w = None
        prep_d = None
        if not tc and chans is not None and w.shape[0] == 5 and stride == 2 and not valid:
            plan_f, plan_d, plan_w = self._tc_plans('tconv5s2' if transposed else 'conv5s2', *chans)
        cache = None
        held = self._planes_of.get(x.data_ptr()) if (tc and scope in self._trunk_row and self._fused()) else None
        assert not getattr(x, '_ic_unwritten', False) or (held is not None and held[0] is x), 'float32 copy of this activation was not written'
        if held is not None and held[0] is x:
            # x exists as unscaled hi/lo planes (written by its batch norm): conv + merge-with-statistics, nothing else
            M = x.numel() // 128
            if self._bn_partial is None or self._bn_partial.numel() * 8 < _lib.lib().ic_nn_bn_partial_bytes(M):
                self._bn_partial = nn.bn_partial_buffer(M, x.device)
            row = self._trunk_row[scope]
            wscale = self._trunk_scales[row]
            if self._prep is None or self._prep_w != x.shape[2]:      # all trunk weights packed once per pass
                self._prep = nn.pack3x3_all(self.groups['ae_w'].w, self._trunk_off, self._trunk_scales, x.shape[2], self._prep_buf)
                self._prep_buf, self._prep_w = self._prep, x.shape[2]
            prep_d = self._prep[row, 1]
            y = nn.conv3x3_tc_fused(held[1], w, wscale, tuple(x.shape), self._bn_partial, prepared=self._prep[row, 0])
            cache = (held[1], wscale)
            self._stats_of = y
            if self.tape is None:           # forward only: nothing reads these planes again
                del self._planes_of[x.data_ptr()]
        elif tc and self.tape is not None:        # keep the input's fp16 planes and the scales for the filter gradient
            y, cache = nn.conv3x3_tc(x, w, keep=True)
        elif plan_f is not None:
            y = plan_f.run(x, w)
        else:
            y = nn.conv3x3_tc(x, w) if tc else nn.conv2d_fwd(x, w, stride, transposed, valid)
        if self.tape is not None:
            tape = self.tape

            def bwd():
                dy = tape.pop(y)
                if dy is None:
                    return
                if tc:        # both gradients on tensor cores (filter gradient: GEMM over pixels, csrc/train_tc.cu)
                    if isinstance(dy, nn.PlanesGrad):       # the batch norm's backward wrote the conv's operand directly
                        # the gradient x already has (residual path) is added in the conv's own output pass
                        had = tape.take(x) if need_dx else None
                        dx, _ = nn.conv3x3_tc_bwd_planes(dy, w, need_dx=need_dx, dw_out=gw, cache=cache, dx_add=had, prepared=prep_d)
                    else:
                        dx, _ = nn.conv3x3_tc_bwd(x, dy, w, need_dx=need_dx, dw_out=gw, cache=cache)
                    if need_dx:
                        tape.acc(x, dx, owned=True)
                    return
                if plan_w is not None:
                    plan_w.run(x, dy, gw)
                else:
                    nn.conv2d_bwd_filter(x, dy, w.shape, stride, transposed, valid, out=gw)
                if mask is not None:
                    nn.mul(gw, mask, out=gw)
                if need_dx:
                    dx = plan_d.run(dy, w) if plan_d is not None else nn.conv2d_bwd_data(dy, w, x.shape, stride, transposed, valid)
                    tape.acc(x, dx, owned=True)
            tape.add(bwd)
        return y

    def _bn(self, x, scope, relu, res1=None, res2=None, f32_out=True):
        gamma, beta = self._w(scope + '/BatchNorm/gamma'), self._w(scope + '/BatchNorm/beta')
        mm = self.stats.view(self.stats.w, scope + '/BatchNorm/moving_mean')
        mv = self.stats.view(self.stats.w, scope + '/BatchNorm/moving_variance')
        planes = None
        if self._fused() and x.dim() == 4 and x.shape[-1] == 128:
            planes = torch.empty(2 * x.numel(), dtype=torch.float16, device=x.device)      # for the 3x3 conv that reads `out`
        partial = self._bn_partial if self._stats_of is x else None
        # x is the output of a fused trunk conv: this batch norm is its only consumer, so in the backward pass dx can go to
        # that conv as pre-scaled fp16 planes (no float32 dx, no maximum search, no split pass)
        grad_as_planes = partial is not None and self.is_training and os.environ.get('IC_TRAIN_FUSED_BWD', '1') != '0'
        self._stats_of = None
        write_out = f32_out or planes is None      # f32_out=False: `out` only feeds the next 3x3 conv, which reads the planes
        if self.is_training:
            upd = self.update_moving
            out, mean, invstd = nn.bn_train_fwd(x, gamma, beta, relu, res1, res2, mm if upd else None, mv if upd else None,
                                                partial=partial, planes_out=planes, write_out=write_out)
        else:           # moving statistics (code/autoencoder.py:115-125 with is_training=False)
            mean, invstd = mm, torch.rsqrt(mv + nn.BN_EPS)
            out, _, _ = nn.bn_train_fwd(x, gamma, beta, relu, res1, res2, stats=(mean, invstd), planes_out=planes, write_out=write_out)
        if planes is not None:
            self._planes_of[out.data_ptr()] = (out, planes)       # holds `out`: its address cannot be reused within the step
        if not write_out:
            out._ic_unwritten = True                # placeholder: only its planes exist
        if self.tape is not None:
            tape, training = self.tape, self.is_training

            def bwd():
                dy = tape.pop(out)
                if dy is None:
                    return
                if grad_as_planes:
                    dx, _, _ = nn.bn_train_bwd_planes(x, dy, gamma, beta, mean, invstd, relu, dgamma=self._g(scope + '/BatchNorm/gamma'),
                                                      dbeta=self._g(scope + '/BatchNorm/beta'))
                    assert id(x) not in tape.grads
                else:
                    dx, _, _ = nn.bn_train_bwd(x, dy, gamma, beta, mean, invstd, relu, use_stats=training,
                                               dgamma=self._g(scope + '/BatchNorm/gamma'), dbeta=self._g(scope + '/BatchNorm/beta'))
                tape.acc(x, dx, owned=True)
                if res1 is not None:
                    tape.acc(res1, dy)
                if res2 is not None:
                    tape.acc(res2, dy)
            tape.add(bwd)
        return out

    def _slim_conv(self, x, scope, relu, res1=None, res2=None, need_dx=True, f32_out=True):
        """slim.conv2d / conv2d_transpose under _batch_norm_scope: conv (no bias) -> BN -> activation (+ fused adds)"""
        y = self._conv(x, self._w(scope + '/weights'), self._g(scope + '/weights'), self._stride[scope], self._tr[scope],
                       need_dx=need_dx, chans=self._shape[scope][1:], scope=scope)
        return self._bn(y, scope, relu, res1, res2, f32_out=f32_out)

    def _res_stack(self, net, prefix, tag, final_scope):
        """15 residual blocks with a skip every 3, a final block without ReLU and the long skip
        (code/autoencoder.py:225-234,253-262,274-287); the adds are fused into the second conv's BN kernel."""
        r0 = net
        for b in range(self.B):
            rb = net
            for i in (1, 2, 3):
                s = '{}/res_block_{}_{}/{}_{}_{}'.format(prefix, tag, b, tag, b, i)
                y = self._slim_conv(net, s + '/conv1', True, f32_out=False)      # conv1's output feeds conv2 only
                net = self._slim_conv(y, s + '/conv2', False, res1=net, res2=rb if i == 3 else None)
        y = self._slim_conv(net, prefix + '/' + final_scope + '/conv1', False, f32_out=False)
        return self._slim_conv(y, prefix + '/' + final_scope + '/conv2', False, res1=net, res2=r0)

    def _encode(self, x):
        E = 'autoencoder/encoder'
        net = nn.normalize_fwd(x)
        net = self._slim_conv(net, E + '/h1', True, need_dx=False)
        net = self._slim_conv(net, E + '/h2', True)
        net = self._res_stack(net, E, 'enc', 'res_block_enc_final')
        bn = self._slim_conv(net, E + '/to_bn', False)
        centers = self._w(E + '/centers')
        enc = nn.hq_fwd(bn, self.C, self.heatmap, centers)
        enc['bn'] = bn
        return enc

    def _decode(self, q_nhwc):
        D = 'autoencoder/decoder'
        net = self._slim_conv(q_nhwc, D + '/from_bn', True)
        net = self._res_stack(net, D, 'dec', 'dec_after_res')
        net = self._slim_conv(net, D + '/h12', True)
        v = self._slim_conv(net, D + '/h13', False)
        x_out = nn.denorm_clip_fwd(v)
        if self.tape is not None:
            tape = self.tape

            def bwd():
                d = tape.pop(x_out)
                if d is not None:
                    tape.acc(v, nn.denorm_clip_bwd(v, d), owned=True)
            tape.add(bwd)
        return x_out

    def _conv3d(self, x, scope, need_dx=True):
        """masked (2,3,3) VALID conv3d (code/probclass.py:227-261) on the depth-major volume x (D, N, H, W, Cin):
        two VALID conv2d passes (filter depth 0 over slices [0, D-1), depth 1 over [1, D)), accumulated."""
        w, gw, mask = self._w(scope + '/weights'), self._g(scope + '/weights'), self._const(scope + '/mask')
        weff = nn.mul(w, mask)
        D, N, H, W, Ci = x.shape
        xa, xb = x[:D - 1].reshape(-1, H, W, Ci), x[1:].reshape(-1, H, W, Ci)
        # layers 1-3 ("other" mask, 24 input channels): both depth passes as ONE tcgen05 conv, forward and data gradient
        plan_f, plan_d, plan_w = self._tc_plans('pc', *self._pc_chans[scope]) if scope != PC_LAYERS[0][0] else (None, None, None)
        if plan_f is not None:
            y = plan_f.run(x, weff)
        else:
            y = nn.conv2d_fwd(xa, weff[0], valid=True)
            y = nn.axpby(1.0, y, 1.0, nn.conv2d_fwd(xb, weff[1], valid=True), out=y)
            y = y.view(D - 1, N, H - 2, W - 2, w.shape[-1])
        if self.tape is not None:
            tape = self.tape

            def bwd():
                dy = tape.pop(y)
                if dy is None:
                    return
                dy4 = dy.reshape(-1, H - 2, W - 2, w.shape[-1])
                if plan_w is not None:          # masked taps come out as zeros
                    plan_w.run(x, dy.contiguous(), gw)
                else:
                    nn.conv2d_bwd_filter(xa, dy4, weff[0].shape, valid=True, out=gw[0])
                    nn.conv2d_bwd_filter(xb, dy4, weff[1].shape, valid=True, out=gw[1])
                    nn.mul(gw, mask, out=gw)
                if need_dx and plan_d is not None:
                    tape.acc(x, plan_d.run(dy.contiguous(), weff), owned=True)
                elif need_dx:
                    dx = torch.empty_like(x)
                    nn.conv2d_bwd_data(dy4, weff[0], xa.shape, valid=True, out=dx[:D - 1].reshape(-1, H, W, Ci))
                    db = nn.conv2d_bwd_data(dy4, weff[1], xb.shape, valid=True).view(D - 1, N, H, W, Ci)
                    if D > 2:
                        nn.axpby(1.0, dx[1:D - 1], 1.0, db[:D - 2], out=dx[1:D - 1])
                    nn.axpby(1.0, db[D - 2], out=dx[D - 1])
                    tape.acc(x, dx, owned=True)
            tape.add(bwd)
        return y

    def _bias_act(self, x, scope, relu, res1=None):
        """tf.nn.bias_add + activation (code/probclass.py:259-261), + the residual add of the 3-D block"""
        C = x.shape[-1]
        bias, ones, zeros = self._w(scope + '/biases'), self._const('ones', C), self._const('zeros', C)
        out, _, _ = nn.bn_train_fwd(x, ones, bias, relu, res1, None, stats=(zeros, ones))
        if self.tape is not None:
            tape = self.tape

            def bwd():
                dy = tape.pop(out)
                if dy is None:
                    return
                dx, _, _ = nn.bn_train_bwd(x, dy, ones, bias, zeros, ones, relu, use_stats=False, dbeta=self._g(scope + '/biases'))
                tape.acc(x, dx, owned=True)
                if res1 is not None:
                    tape.acc(res1, dy)
            tape.add(bwd)
        return out

    def _pc_logits(self, q_nchw, pad_value):
        """pad_for_probclass3d + _ResShallow._logits (code/probclass.py:214-221,268-292) -> logits (C, N, h, w, ceil4(L))"""
        S = [s for s, _ in PC_LAYERS]
        x0 = nn.pc_pad_fwd(q_nchw, pad_value)
        a0 = self._bias_act(self._conv3d(x0, S[0], need_dx=False), S[0], True)
        a1 = self._bias_act(self._conv3d(a0, S[1]), S[1], True)
        c2 = self._conv3d(a1, S[2])
        skip_src = a0[2:]
        skip = nn.crop_fwd(skip_src, 2)                                    # x[:, 2:, 2:-2, 2:-2, :]  (:185-196)
        if self.tape is not None:
            tape = self.tape

            def bwd():
                d = tape.pop(skip)
                if d is not None:
                    g0 = torch.zeros_like(a0)
                    nn.crop_bwd_add(d, g0[2:], 2)
                    tape.acc(a0, g0, owned=True)
            tape.add(bwd)
        net = self._bias_act(c2, S[2], False, res1=skip)
        return self._bias_act(self._conv3d(net, S[3]), S[3], True)

    # ------------------------------------------------------------------ forward-only entry points (autoencoder.py mirror)
    def encode_forward(self, x, is_training=True, update_moving=False):
        """ae.encode(x, is_training) of code/train.py:101 without a tape: batch-statistics batch norm when is_training.
        -> dict(qbar, qhard, qsoft, symbols, z, heatmap) NCHW"""
        assert x.is_cuda and x.dim() == 4 and x.shape[1] == 3 and x.shape[2] % 8 == 0 and x.shape[3] % 8 == 0
        self.is_training, self.update_moving, self.tape = is_training, update_moving, None
        self._begin_pass()
        return self._encode(x.contiguous())

    def decode_forward(self, q, is_training=True, update_moving=False):
        """ae.decode(q, is_training) of code/train.py:102; q NCHW float32 -> x_out NCHW float32 clipped to [0, 255]"""
        self.is_training, self.update_moving, self.tape = is_training, update_moving, None
        self._begin_pass()
        return self._decode(nn.nchw_to_nhwc(q.contiguous().float()))

    # ------------------------------------------------------------------ the graph of code/train.py:86-132
    def _enqueue(self, x, is_training, update_moving, backward):
        """Enqueues forward (+ backward) without any host read-back: the loss scalars stay in self._sums / self._val, the
        rate hinge coefficient is computed on the device.  -> dict of tensors"""
        cfg = self.ae_config
        assert x.is_cuda and x.dim() == 4 and x.shape[1] == 3
        x = x.contiguous()
        N, _, H, W = x.shape
        assert H % 8 == 0 and W % 8 == 0
        self.is_training, self.update_moving = is_training, update_moving
        self.tape = tape = _Tape() if backward else None
        self._begin_pass()
        centers = self._w('autoencoder/encoder/centers')
        # pc.auto_pad_value(ae) = centers[0] (code/probclass.py:59-61), read on the device
        pad_value = centers if self.pc_config.use_centers_for_padding else 0.0
        enc = self._encode(x)
        n_enc = len(tape.fns) if backward else 0
        q_nhwc = nn.nchw_to_nhwc(enc['qbar'])
        x_out = self._decode(q_nhwc)
        logits = self._pc_logits(enc['qbar'], pad_value)                 # stop_gradient(qbar): no backward into q
        bc = nn.pc_xent_fwd(logits, self.L, enc['symbols'])
        # ---- losses (code/train.py:303-336, 352-431): sums of bc, bc * heatmap, w^2 per regularised group
        xf = x if x.dtype == torch.float32 else x.float()
        Lb = _lib.lib()
        sums, ws = self._sums, self._loss_ws
        hm = enc['heatmap']

        def masked_sums(a, b, n, out):
            _lib.check(Lb.ic_masked_sums_fwd(_lib.ptr(a), _lib.ptr(b), n, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        masked_sums(bc, hm, bc.numel(), sums)
        for i, g in enumerate((self.groups['ae_w'], self.groups['pc_w'], self.groups['ae_centers'])):
            masked_sums(g.w, g.w, g.w.numel(), sums[2 + 2 * i:])
        if cfg.distortion_to_minimize == 'ms_ssim':
            d_xout, self._val = nn.msssim_tf_bwd(xf, x_out, -float(cfg.K_ms_ssim))    # d/dx_out of K (1 - MS-SSIM)
        else:       # 'mse' / 'psnr' (code/train.py:381-397): self._val = per-image float MSE
            d_xout, self._val = nn.distortion_bwd(xf, x_out, psnr=cfg.distortion_to_minimize == 'psnr')
        tensors = dict(bc=bc, heatmap=hm, x_out=x_out, symbols=enc['symbols'], qbar=enc['qbar'], z=enc['z'])
        if not backward:
            return tensors
        # ---- backward, in reverse order of the graph
        n = bc.numel()
        coef = nn.rate_coef(sums, n, cfg.beta, cfg.H_target, hm is not None, self._coef)
        tape.acc(logits, nn.pc_xent_bwd(logits, self.L, enc['symbols'], hm, coef_dev=coef), owned=True)
        tape.acc(x_out, d_xout, owned=True)
        bn = enc['bn']

        def hq_bwd():
            dq = tape.pop(q_nhwc)
            dhm = nn.scale_dev(coef, bc) if hm is not None else None                  # d pc_loss / d heatmap3D
            dbn, _ = nn.hq_bwd(bn, self.C, self.heatmap, centers, dq, dhm, dcenters=self._g('autoencoder/encoder/centers'))
            tape.acc(bn, dbn, owned=True)
        # the tape so far: encoder ops, decoder ops, context-model ops.  The quantizer's backward must run after the
        # decoder's and before the encoder's: insert it where the encoder ends.
        tape.fns.insert(n_enc, hq_bwd)
        tape.backward()
        self.tape = None
        self._planes_of = {}
        return tensors

    def _read_losses(self, shape, n_symbols, tensors):
        """the step's one read-back: loss scalars -> the components code/train.py logs"""
        cfg = self.ae_config
        N, _, H, W = shape
        host = self._sums.cpu().tolist()
        if cfg.distortion_to_minimize == 'ms_ssim':
            msssim = float(self._val.item())
            d_loss = cfg.K_ms_ssim * (1.0 - msssim)
        else:
            msssim = None
            mse = self._val.double().cpu().numpy()
            d_loss = float(mse.mean()) if cfg.distortion_to_minimize == 'mse' else cfg.K_psnr - float((10 * np.log10(255.0 * 255.0 / mse)).mean())
        H_real = host[0] / n_symbols
        H_mask = host[1] / n_symbols if self.heatmap else H_real
        H_soft = 0.5 * (H_mask + H_real)
        pc_loss = cfg.beta * max(H_soft - cfg.H_target, 0.0)
        reg_enc_dec = cfg.regularization_factor * 0.5 * host[3]
        if cfg.regularization_factor_centers != 0:
            reg_enc_dec += cfg.regularization_factor_centers * 0.5 * host[7]
        rf_pc = self.pc_config.regularization_factor
        reg_pc = 0.0 if rf_pc is None else rf_pc * 0.5 * host[5]
        return dict(total_loss=d_loss + pc_loss + reg_enc_dec + reg_pc, d_loss_scaled=d_loss, pc_loss=pc_loss, H_real=H_real,
                    H_mask=H_mask, ms_ssim=msssim, reg=reg_enc_dec + reg_pc, bpp=host[0] / (N * H * W), tensors=tensors)

    def forward_backward(self, x, is_training=True, update_moving=False, backward=True):
        """-> dict of loss components (floats) and tensors; gradients are left in the groups' g buffers."""
        tensors = self._enqueue(x, is_training, update_moving, backward)
        return self._read_losses(x.shape, tensors['bc'].numel(), tensors)

    # ------------------------------------------------------------------ optimiser (code/train.py:339-349)
    def learning_rates(self):
        """(lr_ae, lr_pc) at the current global step"""
        return (learning_rate_at(self.ae_config, self.global_step, self.num_itr_per_epoch),
                learning_rate_at(self.pc_config, self.global_step, self.num_itr_per_epoch))

    def _adam_plan(self):
        cfg = self.ae_config
        plan = []
        if cfg.train_autoencoder:
            plan += [('ae_w', 0, cfg.regularization_factor), ('ae_bn', 0, 0.0), ('ae_centers', 0, cfg.regularization_factor_centers)]
        if cfg.train_probclass:
            plan += [('pc_w', 1, self.pc_config.regularization_factor or 0.0), ('pc_b', 1, 0.0)]
        return plan

    def _stage_step_sizes(self):
        """tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), t counting from 1; written to device memory
        (pinned staging buffer, asynchronous copy) so that the Adam launches do not change from step to step"""
        self._adam_t += 1
        t = self._adam_t
        corr = math.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
        lr_ae, lr_pc = self.learning_rates()
        self._lr_host[0], self._lr_host[1] = lr_ae * corr, lr_pc * corr
        self._lr_dev.copy_(self._lr_host, non_blocking=True)

    def _enqueue_adam(self):
        for name, which, l2 in self._adam_plan():
            g = self.groups[name]
            nn.adam_step_dev(g.w, g.g, g.m, g.v, self._lr_dev[which:], l2=l2)

    def apply_gradients(self):
        self._stage_step_sizes()
        self._enqueue_adam()
        self.global_step += 1

    def step(self, x):
        """one sess.run(train_op) of the reference: forward, loss, backward, BN moving averages, both Adam updates.
        With enable_cuda_graph() the whole step (~1250 kernel launches) is one graph replay."""
        if self._graph is not None and tuple(x.shape) == tuple(self._x_static.shape) and x.dtype == self._x_static.dtype:
            self._x_static.copy_(x, non_blocking=True)
            self._stage_step_sizes()
            self._graph.replay()
            self.global_step += 1
            self._val = self._graph_val          # the distortion scalar the graph writes (an eager call in between re-binds _val)
            return self._read_losses(x.shape, self._graph_tensors['bc'].numel(), self._graph_tensors)
        tensors = self._enqueue(x, True, True, True)
        self.apply_gradients()
        return self._read_losses(x.shape, tensors['bc'].numel(), tensors)

    def enable_cuda_graph(self, x_example):
        """Captures forward + loss + backward + moving averages + Adam for batches of x_example's shape / dtype in ONE CUDA
        graph.  Nothing in the step reads back to the host (the hinge coefficient, the pad value and the Adam step sizes
        live in device memory), so a replay enqueues ~1250 kernels without per-launch host work.  The tensors of the
        returned dicts are then views into graph memory: valid until the next step."""
        self._x_static = x_example.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                # warm-up: first-call initialisations, workspace growth (no state change)
            for _ in range(2):
                self._enqueue(self._x_static, True, False, True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._graph_tensors = self._enqueue(self._x_static, True, True, True)
            self._enqueue_adam()
        self._graph = graph
        # the captured launches hold raw pointers into the shared scratch buffer of nn.py: keep THAT tensor alive for as
        # long as the graph is, whatever a later, larger call elsewhere replaces the shared one with
        self._graph_ws = nn.current_workspace()
        # same for the batch-norm partial-sum buffer of the fused trunk (a later, larger eager call replaces self._bn_partial)
        self._graph_partial = self._bn_partial
        self._graph_val = self._val
        return self


def train_loop(trainer, batches, log_interval=100, log=print):
    """code/train.py:216-269 without the TF session / summary plumbing: runs trainer.step over an iterable of batches"""
    last = None
    for x in batches:
        last = trainer.step(x)
        itr = trainer.global_step
        if log_interval and itr % log_interval == 0:
            log('{}: loss {:.4f}  d_loss {:.4f}  pc_loss {:.4f}  H_real {:.4f}  bpp {:.4f}  ms_ssim {:.5f}'.format(
                itr, last['total_loss'], last['d_loss_scaled'], last['pc_loss'], last['H_real'], last['bpp'], last['ms_ssim']))
    return last


def random_crops(images, batch_size, crop, rng):
    """What code/inputpipeline.py:147-213 feeds the graph: `batch_size` random crop_size crops with a random horizontal flip,
    NCHW uint8.  images: list of H x W x 3 (or 3 x H x W) uint8 arrays at least as large as the crop."""
    ch, cw = crop
    out = np.empty((batch_size, 3, ch, cw), np.uint8)
    for i in range(batch_size):
        im = images[rng.randint(len(images))]
        if im.shape[0] == 3 and im.ndim == 3 and im.shape[2] != 3:
            im = im.transpose(1, 2, 0)
        y, x = rng.randint(im.shape[0] - ch + 1), rng.randint(im.shape[1] - cw + 1)
        c = im[y:y + ch, x:x + cw, :3]
        if rng.randint(2):
            c = c[:, ::-1]
        out[i] = c.transpose(2, 0, 1)
    return out


def main():
    """train.py-style run (code/train.py:471-527 without the TF session / checkpoint / TensorBoard plumbing):

        python -m imgcomp_cvpr_b200.trainer --ae_config cvpr/med --steps 200 [--images 'train/*.png'] [--save out.npz]
    """
    import argparse
    import glob
    import time
    from . import config
    ap = argparse.ArgumentParser(description='training steps of code/train.py on one B200')
    ap.add_argument('--ae_config', default='cvpr/med')
    ap.add_argument('--pc_config', default='cvpr/res_shallow')
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--batch_size', type=int, default=None, help='default: the config value (30)')
    ap.add_argument('--images', default=None, help='glob of training images (default: seeded synthetic images)')
    ap.add_argument('--weights', default=None, help='.npz of TF variable name -> array, or a TensorFlow checkpoint prefix / directory, to start from (default: seeded synthetic weights)')
    ap.add_argument('--save', default=None, help='write the trained variables (.npz, TF names) here')
    ap.add_argument('--save_state', default=None, help='write variables + Adam moments + step counters (.npz) here: --resume input')
    ap.add_argument('--resume', default=None, help='.npz written by --save_state (the reference: --restore, code/train.py:487-499)')
    ap.add_argument('--num_itr_per_epoch', type=int, default=1000, help='for the staircase LR decay (training_helpers.py:51-60)')
    ap.add_argument('--log_interval', type=int, default=10)
    ap.add_argument('--mode', default='exact', choices=['fp32', 'exact'])
    ap.add_argument('--no_graph', action='store_true', help='do not capture the step in a CUDA graph')
    ap.add_argument('--seed', type=int, default=0)
    args = ap.parse_args()
    a, p = config.ae_config(args.ae_config), config.pc_config(args.pc_config)
    B = args.batch_size or a.batch_size
    from . import tf_checkpoint
    W = tf_checkpoint.load_weights(args.weights) if args.weights else weights_mod.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k,
                                                                                       a.arch_param_B)
    rng = np.random.RandomState(args.seed)
    if args.images:
        from PIL import Image
        images = [np.asarray(Image.open(f).convert('RGB')) for f in sorted(glob.glob(args.images))]
        images = [im for im in images if im.shape[0] >= a.crop_size[0] and im.shape[1] >= a.crop_size[1]]
        assert images, 'no image at least as large as the crop %s' % (a.crop_size,)
    else:
        images = list(weights_mod.synthetic_images(64, 2 * a.crop_size[0], 2 * a.crop_size[1], seed=args.seed + 1))
    tr = Trainer(a, p, W, num_itr_per_epoch=args.num_itr_per_epoch, mode=args.mode)
    if args.resume:
        tr.load_state_dict(dict(np.load(args.resume)))
        print('resumed at global step', tr.global_step)
    pinned = torch.empty((B, 3) + tuple(a.crop_size), dtype=torch.uint8).pin_memory()
    x = torch.empty_like(pinned, device='cuda')
    if not args.no_graph:
        pinned.copy_(torch.from_numpy(random_crops(images, B, a.crop_size, rng)))
        tr.enable_cuda_graph(x.copy_(pinned))
    t0, n0 = time.perf_counter(), 0
    for itr in range(1, args.steps + 1):
        pinned.copy_(torch.from_numpy(random_crops(images, B, a.crop_size, rng)))
        out = tr.step(x.copy_(pinned, non_blocking=True))
        if itr % args.log_interval == 0 or itr == args.steps:
            dt = time.perf_counter() - t0
            print('{:6d}  loss {:9.3f}  d_loss {:9.3f}  pc_loss {:8.3f}  H_real {:.4f}  H_mask {:.4f}  bpp {:.4f}  ms_ssim {:.5f}'
                  '  (img/s: {:.1f})'.format(itr, out['total_loss'], out['d_loss_scaled'], out['pc_loss'], out['H_real'], out['H_mask'],
                                             out['bpp'], out['ms_ssim'], (itr - n0) * B / dt), flush=True)
            t0, n0 = time.perf_counter(), itr
    if args.save:
        np.savez(args.save, **tr.weights())
        print('saved', args.save)
    if args.save_state:
        np.savez(args.save_state, **tr.state_dict())
        print('saved state', args.save_state)


if __name__ == '__main__':
    main()
