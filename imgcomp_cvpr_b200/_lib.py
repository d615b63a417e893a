"""ctypes binding of libimgcomp_b200.so (include/imgcomp_b200.h).

There is no CPU fallback: if the shared object is missing or no B200-class
device is usable, every entry point raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_longlong, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libimgcomp_b200.so')

IC_MODE_FP32, IC_MODE_EXACT, IC_MODE_FAST = 0, 1, 2
MODES = {'fp32': IC_MODE_FP32, 'exact': IC_MODE_EXACT, 'fast': IC_MODE_FAST}


class IcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('imgcomp_b200 error %d: %s' % (code, msg))
        self.code = code


class AeConfig(ctypes.Structure):
    _fields_ = [('num_chan_bn', c_int32), ('arch_param_B', c_int32), ('num_centers', c_int32),
                ('heatmap', c_int32), ('normalization', c_int32)]


class PcConfig(ctypes.Structure):
    _fields_ = [('kernel_size', c_int32), ('arch_param_k', c_int32), ('num_centers', c_int32)]


# name -> (restype, argtypes): every symbol include/imgcomp_b200.h declares
SIGNATURES = {
    'ic_last_error': (c_char_p, []),
    'ic_abi_version': (c_int, []),
    'ic_device_ok': (c_int, []),
    'ic_ae_num_tensors': (c_int, [POINTER(AeConfig)]),
    'ic_ae_tensor_name': (c_char_p, [POINTER(AeConfig), c_int]),
    'ic_ae_tensor_numel': (c_int64, [POINTER(AeConfig), c_int]),
    'ic_ae_create': (c_int, [POINTER(AeConfig), POINTER(c_void_p), c_int, POINTER(c_void_p)]),
    'ic_ae_destroy': (None, [c_void_p]),
    'ic_encode_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    'ic_encode_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 7 +
                      [c_void_p, c_size_t, c_int, c_void_p]),
    'ic_decode_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    'ic_decode_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_int, c_void_p]),
    'ic_ae_centers': (c_int, [c_void_p, c_void_p, c_void_p]),
    'ic_quantize_fwd': (c_int, [c_void_p, c_void_p, c_int, c_float, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ic_pc_num_tensors': (c_int, [POINTER(PcConfig)]),
    'ic_pc_tensor_name': (c_char_p, [POINTER(PcConfig), c_int]),
    'ic_pc_tensor_numel': (c_int64, [POINTER(PcConfig), c_int]),
    'ic_pc_create': (c_int, [POINTER(PcConfig), POINTER(c_void_p), c_int, POINTER(c_void_p)]),
    'ic_pc_destroy': (None, [c_void_p]),
    'ic_pc_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    'ic_pc_bitcost_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_pc_logits_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_pc_freqs_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_pc_codec_freqs_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_pc_codec_freqs_u32_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                          c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_pc_decode_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    'ic_pc_decode_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_pc_context_freqs_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_conv2d_workspace_bytes': (c_size_t, [c_int] * 10),
    'ic_nn_conv2d_fwd': (c_int, [c_void_p, c_void_p] + [c_int] * 10 + [c_void_p, c_void_p]),
    'ic_nn_conv2d_bwd_data': (c_int, [c_void_p, c_void_p] + [c_int] * 10 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_conv2d_bwd_filter': (c_int, [c_void_p, c_void_p] + [c_int] * 10 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_conv3x3_tc_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'ic_nn_conv3x3_tc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_conv3x3_tc_ex': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                    c_void_p]),
    'ic_nn_conv3x3_tc_bwd_ex': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    'ic_nn_conv3x3_tc_bwd_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'ic_nn_conv3x3_tc_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_tc_plan_create': (c_int, [c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    'ic_nn_tc_plan_destroy': (None, [c_void_p]),
    'ic_nn_tc_plan_map': (c_int64, [c_void_p, c_void_p, c_int64]),
    'ic_nn_tc_plan_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    'ic_nn_tc_plan_run': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_tc_wgrad_plan_create': (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    'ic_nn_tc_wgrad_plan_destroy': (None, [c_void_p]),
    'ic_nn_tc_wgrad_plan_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    'ic_nn_tc_wgrad_plan_run': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_bn_workspace_bytes': (c_size_t, [c_int64, c_int]),
    'ic_nn_bn_train_fwd': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_weight_scales': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    'ic_nn_conv3x3_tc_fused_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'ic_nn_bn_partial_bytes': (c_size_t, [c_int64]),
    'ic_nn_conv3x3_tc_prepared_bytes': (c_size_t, []),
    'ic_nn_pack3x3_all': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'ic_nn_conv3x3_tc_fused': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    'ic_nn_bn_train_fwd_ex': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t,
                                      c_void_p]),
    'ic_nn_bn_train_bwd': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_bn_train_bwd_ex': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    'ic_nn_conv3x3_tc_bwd_planes': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_hq_workspace_bytes': (c_size_t, [c_int64]),
    'ic_nn_hq_bwd': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_nn_denorm_clip_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ic_nn_denorm_clip_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ic_nn_nhwc_to_nchw': (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p]),
    'ic_nn_nchw_to_nhwc': (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p]),
    'ic_nn_axpby': (c_int, [c_float, c_void_p, c_float, c_void_p, c_int64, c_void_p, c_void_p]),
    'ic_nn_mul': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'ic_nn_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int64,
                                c_float, c_void_p, c_void_p]),
    'ic_nn_normalize_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ic_nn_hq_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int] + [c_void_p] * 7),
    'ic_nn_pc_pad_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    'ic_nn_pc_xent_fwd': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ic_nn_pc_xent_bwd': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                  c_void_p, c_void_p, c_void_p]),
    'ic_nn_rate_coef': (c_int, [c_void_p, c_int64, c_float, c_float, c_int, c_void_p, c_void_p]),
    'ic_nn_scale_dev': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'ic_nn_adam_step_dev': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float, c_float,
                                    c_void_p, c_void_p]),
    'ic_nn_crop_fwd': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ic_nn_crop_bwd_add': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ic_msssim_bwd_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'ic_msssim_tf_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_msssim_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'ic_msssim_tf_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_msssim_np_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_loss_workspace_bytes': (c_size_t, []),
    'ic_masked_sums_fwd': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ic_mse_per_image_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    'ic_nn_distortion_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    'ic_debug_conv3x3': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_size_t, c_int, c_void_p]),
    'ic_launch_count': (c_longlong, []),
    'ic_profile_enable': (None, [c_int]),
    'ic_profile_reset': (None, []),
    'ic_profile_get': (c_int, [c_int, POINTER(c_double), POINTER(c_longlong)]),
    'ic_crc32c': (ctypes.c_uint32, [c_void_p, c_int64]),
    'ic_ac_enc_create': (c_int, [POINTER(c_void_p)]),
    'ic_ac_enc_write': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64]),
    'ic_ac_enc_write_u32': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64]),
    'ic_ac_enc_finish': (c_int, [c_void_p, POINTER(POINTER(c_uint8)), POINTER(c_int64), POINTER(c_int64)]),
    'ic_ac_enc_destroy': (None, [c_void_p]),
    'ic_ac_dec_create': (c_int, [c_void_p, c_int64, POINTER(c_void_p)]),
    'ic_ac_dec_read': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64]),
    'ic_ac_dec_destroy': (None, [c_void_p]),
}

_lib = None


def lib():
    """The loaded shared object (loads on first use; raises if it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError('%s not found: build it with `python -m imgcomp_cvpr_b200.build` '
                              '(there is no CPU fallback)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise IcError(rc, lib().ic_last_error().decode())


def require_device():
    if not lib().ic_device_ok():
        raise IcError(-2, 'no usable sm_100 CUDA device (imgcomp_b200 has no CPU path)')


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def host_tensor_array(arrays):
    """list of contiguous float32 numpy arrays -> (POINTER(c_void_p) array, keepalive)."""
    arr = (c_void_p * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data
    return arr
