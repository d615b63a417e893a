"""Host mirror of code/train.py: get_loss (:303-336) and Distortions (:352-431) as forward-only functions on CUDA
tensors, and get_train_op (:339-349) as the constructor of trainer.Trainer, which owns the whole training step
(forward + loss + backward + the two Adam optimisers)."""
import math

import numpy as np
import torch

from . import _lib, ms_ssim


def _l2(w):
    """tf.nn.l2_loss: sum(w^2) / 2"""
    return float(np.sum(np.asarray(w, np.float64) ** 2) / 2)


def regularization_losses(ae_config, pc_config, weights):
    """(reg_enc, reg_dec, reg_probclass) as the slim l2 regularisers of the reference define them:
    regularization_factor * l2_loss(w) for every conv weight of the scope (code/autoencoder.py:101-102),
    plus regularization_factor_centers * l2_loss(centers) in the encoder scope (code/quantizer.py:18-24);
    the probclass term is None unless pc_config.regularization_factor is set (code/probclass.py:90-95,121-125)."""
    f = ae_config.regularization_factor
    enc = sum(f * _l2(w) for k, w in weights.items() if k.startswith('autoencoder/encoder/') and k.endswith('/weights'))
    if ae_config.regularization_factor_centers != 0:
        enc += ae_config.regularization_factor_centers * _l2(weights['autoencoder/encoder/centers'])
    dec = sum(f * _l2(w) for k, w in weights.items() if k.startswith('autoencoder/decoder/') and k.endswith('/weights'))
    pcr = None
    if pc_config.regularization_factor is not None:
        pcr = pc_config.regularization_factor * sum(_l2(w) for k, w in weights.items()
                                                    if k.startswith('probclass3d/') and k.endswith('/weights'))
    return enc, dec, pcr


def get_loss(config, ae, pc, d_loss_scaled, bc, heatmap, reg=None):
    """code/train.py:303-336.  bc, heatmap: NCHW float32 CUDA tensors.  The regularisation terms come from
    pc.regularization_loss() / ae.encoder_regularization_loss() / ae.decoder_regularization_loss() as in the reference
    (:321-326) unless `reg` = regularization_losses(...) is given.
    -> (total_loss, H_real, pc_comps, ae_comps) as Python floats."""
    if reg is None:
        reg = (ae.encoder_regularization_loss(), ae.decoder_regularization_loss(), pc.regularization_loss())
    assert config.H_target
    bc = bc.contiguous().float()
    hm = heatmap.contiguous().float() if heatmap is not None else None
    L = _lib.lib()
    ws = torch.empty(L.ic_loss_workspace_bytes(), dtype=torch.uint8, device=bc.device)
    out = torch.empty(2, dtype=torch.float64, device=bc.device)
    _lib.check(L.ic_masked_sums_fwd(_lib.ptr(bc), _lib.ptr(hm), bc.numel(), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                    _lib.stream_ptr()))
    s_bc, s_mask = out.cpu().tolist()
    H_real = s_bc / bc.numel()
    H_mask = (s_mask / bc.numel()) if hm is not None else H_real
    H_soft = 0.5 * (H_mask + H_real)
    pc_loss = config.beta * max(H_soft - config.H_target, 0.0)
    reg_enc, reg_dec, reg_probclass = reg
    if reg_probclass is None:
        reg_probclass = 0
    reg_loss = reg_probclass + reg_enc + reg_dec
    pc_comps = [('H_mask', H_mask), ('H_real', H_real), ('pc_loss', pc_loss), ('reg', reg_probclass)]
    ae_comps = [('d_loss_scaled', float(d_loss_scaled)), ('reg_enc_dec', reg_enc + reg_dec)]
    total_loss = float(d_loss_scaled) + pc_loss + reg_loss
    return total_loss, H_real, pc_comps, ae_comps


class Distortions(object):
    """code/train.py:352-431, forward only."""

    def __init__(self, config, x, x_out, is_training):
        assert x.dtype == torch.float32 and x_out.dtype == torch.float32
        self.config = config
        minimize_for = config.distortion_to_minimize
        assert minimize_for in ('mse', 'psnr', 'ms_ssim')
        should_get_ms_ssim = minimize_for == 'ms_ssim'
        cast_to_int_for_psnr = (not is_training) or minimize_for != 'psnr'
        cast_to_int_for_mse = (not is_training) or minimize_for != 'mse'
        self.mse = float(self.get_mse_per_img(x, x_out, cast_to_int_for_mse).mean())
        self.psnr = float(self.get_psnr_per_image(x, x_out, cast_to_int_for_psnr).mean())
        self.ms_ssim = float(self.get_ms_ssim(x, x_out)) if should_get_ms_ssim else None
        self.d_loss_scaled = self._get_distortion_to_minimize(minimize_for)

    def _get_distortion_to_minimize(self, minimize_for):
        if minimize_for == 'mse':
            return self.mse
        if minimize_for == 'psnr':
            return self.config.K_psnr - self.psnr
        if minimize_for == 'ms_ssim':
            return self.config.K_ms_ssim * (1 - self.ms_ssim)
        raise ValueError('Invalid: {}'.format(minimize_for))

    @staticmethod
    def get_mse_per_img(inp, otp, cast_to_int):
        """float32 tensor of shape (N,) (code/train.py:400-418)"""
        a, b = inp.contiguous().float(), otp.contiguous().float()
        assert a.shape == b.shape and a.dim() == 4
        out = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
        _lib.check(_lib.lib().ic_mse_per_image_fwd(_lib.ptr(a), _lib.ptr(b), a.shape[0], a[0].numel(), int(bool(cast_to_int)),
                                                   _lib.ptr(out), _lib.stream_ptr()))
        return out

    @staticmethod
    def get_psnr_per_image(inp, otp, cast_to_int):
        """10 * log10(255^2 / mse) (code/train.py:420-425)"""
        mse = Distortions.get_mse_per_img(inp, otp, cast_to_int)
        return 10 * torch.log(255.0 * 255.0 / mse) / math.log(10.0)

    @staticmethod
    def get_ms_ssim(inp, otp):
        return ms_ssim.MultiScaleSSIM(inp, otp, data_format='NCHW', name='MS-SSIM')


def get_train_op(ae_config, pc_config, weights, num_itr_per_epoch=1000):
    """code/train.py:339-349 builds Adam_AE (lr_ae) for the autoencoder variables and Adam_PC (lr_pc) for
    pc.variables() over total_loss.  Eager equivalent: a trainer.Trainer; `op.step(x)` runs one
    sess.run([train_op, global_step]) of train_loop (:252) on the batch x (N,3,H,W uint8/float32 CUDA)."""
    from . import trainer
    return trainer.Trainer(ae_config, pc_config, weights, num_itr_per_epoch)
