"""ctypes binding of libctgan_sm100.so (include/ctgan_sm100.h).

There is NO fallback: if the shared library is missing this module raises at import,
and every entry point raises RuntimeError with `ctgan_last_error()` on a non-zero
return code.
"""
import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_uint64, c_float, c_void_p, c_char_p, POINTER, Structure

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CTGAN_SM100_LIB') or os.path.join(_HERE, 'libctgan_sm100.so')   # override: A/B builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "ctgan_b200: %s not found. Build it with `python -m ctgan_b200.build` "
        "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

F32, BF16 = 0, 1
EPI_RELU = 1
EPI_RES_UP2 = 2
EPI_OUT_S2D = 4
EPI_OUT_D2S = 8
EPI_S2D_SKIP = 256
BN_RELU, BN_UP2, BN_ACCUM = 1, 2, 4


class ConvDesc(Structure):
    _fields_ = [(n, c_int32) for n in (
        'N', 'H', 'W', 'Cin', 'Ho', 'Wo', 'Cout', 'kh', 'kw', 'stride', 'pad_t', 'pad_l', 'x_dtype', 'y_dtype')]


class LossDesc(Structure):
    _fields_ = [('B', c_int32), ('NF', c_int32), ('F', c_int32), ('P', c_int32), ('n_classes', c_int32),
                ('feat_dtype', c_int32), ('lambda_gp', c_float), ('lambda2', c_float), ('factor_m', c_float),
                ('acgan_scale', c_float)]


P = c_void_p
_PROTOS = {
    'ctgan_version': (c_int, []),
    'ctgan_kernel_launches': (ctypes.c_ulonglong, []),
    'ctgan_last_error': (c_char_p, []),
    'ctgan_tc_available': (c_int, []),
    'ctgan_conv_fprop': (c_int, [POINTER(ConvDesc), P, P, P, P, c_int, P]),
    'ctgan_conv_dgrad': (c_int, [POINTER(ConvDesc), P, P, P, P]),
    'ctgan_conv_wgrad': (c_int, [POINTER(ConvDesc), P, P, P, c_int, P]),
    'ctgan_set_fprop_halo': (None, [c_int]),
    'ctgan_set_fprop_variant': (None, [c_int]),
    'ctgan_set_pdl': (None, [c_int]),
    'ctgan_set_splitk': (None, [c_int]),
    'ctgan_set_wgrad_variant': (None, [c_int]),
    'ctgan_conv_fprop_tc': (c_int, [POINTER(ConvDesc), P, P, P, P, P, c_int, P]),
    'ctgan_conv_fprop_tc_masked': (c_int, [POINTER(ConvDesc), P, P, P, P, P, P, c_int, P]),
    'ctgan_conv_fprop_tc_actdrop': (c_int, [POINTER(ConvDesc), P, P, P, P, P, c_float, c_float, c_uint64, c_uint64, P, c_int, P]),
    'ctgan_conv_wgrad_tc': (c_int, [POINTER(ConvDesc), P, P, P, P]),
    'ctgan_conv_wgrad_tc_multi_ok': (c_int, [POINTER(ConvDesc)]),
    'ctgan_set_wgrad_multi_chunk': (None, [c_int]),
    'ctgan_set_wgrad_multi_balance': (None, [c_int, c_int]),
    'ctgan_wgrad_multi_last_gain_pct': (c_int, []),
    'ctgan_wgrad_multi_assign': (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    'ctgan_set_fprop_nores': (None, [c_int]),
    'ctgan_conv_wgrad_tc_multi': (c_int, [c_int, POINTER(ConvDesc), POINTER(P), POINTER(P), POINTER(P), P]),
    'ctgan_conv_wgrad_tc_multi_embed': (c_int, [c_int, POINTER(ConvDesc), POINTER(P), POINTER(P), POINTER(P), POINTER(c_int), P]),
    'ctgan_set_wgrad_multi_items_per_sm': (None, [c_int]),
    'ctgan_conv_tf32_ok': (c_int, [POINTER(ConvDesc)]),
    'ctgan_conv_fprop_tf32': (c_int, [POINTER(ConvDesc), P, P, P, P, P, P, c_int, P]),
    'ctgan_conv_wgrad_tf32_multi_ok': (c_int, [POINTER(ConvDesc)]),
    'ctgan_conv_wgrad_tf32_multi': (c_int, [c_int, POINTER(ConvDesc), POINTER(P), POINTER(P), POINTER(P), P]),
    'ctgan_pack_filter_f32': (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    'ctgan_pack_filters_multi_f32': (c_int, [P, P, P, c_int, P]),
    'ctgan_pack_filter_bf16': (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    'ctgan_pack_filters_multi': (c_int, [P, P, P, c_int, P]),
    'ctgan_im2col_thin': (c_int, [POINTER(ConvDesc), c_int, c_int, P, P, P]),
    'ctgan_col2im_thin': (c_int, [POINTER(ConvDesc), c_int, c_int, P, P, P, P]),
    'ctgan_pack_filter_thin': (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    'ctgan_wgrad_thin_tc': (c_int, [P, P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    'ctgan_space_to_depth': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_depth_to_space': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_space_to_depth_mul': (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_depth_to_space_mul': (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_space_to_depth_mask': (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_box_filter': (c_int, [P, P, c_int, c_int, P]),
    'ctgan_box_filter_grad': (c_int, [P, P, c_int, c_int, P]),
    'ctgan_pack_filter_s2d': (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_s2d_filter_grad': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_im2col_strided': (c_int, [POINTER(ConvDesc), c_int, P, P, P]),
    'ctgan_col2im_strided': (c_int, [POINTER(ConvDesc), c_int, P, P, P, P]),
    'ctgan_pack_filter_padk': (c_int, [P, P, P, c_int, c_int, P]),
    'ctgan_add_prefix': (c_int, [P, P, c_int64, c_int, P]),
    'ctgan_bias_grad': (c_int, [P, P, c_int64, c_int, c_int, c_int, P]),
    'ctgan_bias_add': (c_int, [P, P, P, c_int64, c_int, c_int, P]),
    'ctgan_cast': (c_int, [P, c_int, P, c_int, c_int64, P]),
    'ctgan_add': (c_int, [P, P, P, c_int64, c_int, P]),
    'ctgan_mul': (c_int, [P, P, P, c_int64, c_int, P]),
    'ctgan_scale': (c_int, [P, c_float, P, c_int64, c_int, P]),
    'ctgan_act_dropout_fwd': (c_int, [P, P, P, P, c_int64, c_int, c_float, c_float, c_uint64, c_uint64, P, P]),
    'ctgan_fork_dropout_relu': (c_int, [P, P, P, P, P, P, c_int64, c_int, c_float, c_uint64, c_uint64, P, P]),
    'ctgan_mask_sum2': (c_int, [P, P, P, P, P, c_int64, c_int, P]),
    'ctgan_mask_fork2': (c_int, [P, P, P, P, P, c_int64, c_int, P]),
    'ctgan_mul_relu_mask': (c_int, [P, P, P, c_int64, c_int, P]),
    'ctgan_pool_add_fork': (c_int, [c_int, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, c_uint64, c_uint64, P, P]),
    'ctgan_mask_sum2_up': (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_unary_fwd': (c_int, [P, P, c_int64, c_int, c_int, P]),
    'ctgan_unary_bwd': (c_int, [P, P, P, c_int64, c_int, c_int, P]),
    'ctgan_pool2x2': (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, c_int, P]),
    'ctgan_upsample2x': (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, c_int, P]),
    'ctgan_spatial_sum': (c_int, [P, P, c_int, c_int, c_int, c_float, c_int, P]),
    'ctgan_spatial_bcast': (c_int, [P, P, c_int, c_int, c_int, c_float, c_int, P]),
    'ctgan_nchw_to_nhwc': (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_nhwc_to_nchw': (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_crop': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_crop_bwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_prep_real': (c_int, [P, P, c_int64, c_float, c_float, c_uint64, c_uint64, P, P]),
    'ctgan_prep_real_u8': (c_int, [P, P, c_int64, c_float, c_float, c_uint64, c_uint64, P, P]),
    'ctgan_prep_real_dup': (c_int, [P, c_int, P, P, c_int64, c_float, c_float, c_uint64, c_uint64, P, P]),
    'ctgan_memset_zero': (c_int, [P, c_int64, P]),
    'ctgan_interpolate': (c_int, [P, P, P, P, c_int, c_int, P]),
    'ctgan_bn_workspace_floats': (c_int64, [c_int, c_int, c_int, c_int]),
    'ctgan_bn_fwd': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_int, c_int, c_int, P]),
    'ctgan_bn_bwd': (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_bn_fused_ok': (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    'ctgan_bn_fwd_fused': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, c_int, c_int, P]),
    'ctgan_bn_bwd_fused': (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'ctgan_ln_workspace_floats': (c_int64, [c_int, c_int64]),
    'ctgan_ln_fwd': (c_int, [P, P, P, P, P, P, P, c_int, c_int64, c_int, c_float, c_int, P]),
    'ctgan_ln_core': (c_int, [P, P, P, P, P, P, P, c_int, c_int64, c_int, c_int, c_int, c_int, P]),
    'ctgan_ln_param_grad': (c_int, [P, P, P, P, P, P, c_int, c_int64, c_int, c_int, P]),
    'ctgan_ln_bwd2_x': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int64, c_int, c_int, P]),
    'ctgan_ct_gp_loss_fwd': (c_int, [POINTER(LossDesc), P, P, P, P, P, P, P, P, P, P, P]),
    'ctgan_ct_gp_loss_bwd': (c_int, [POINTER(LossDesc), P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P]),
    'ctgan_mean_fwd': (c_int, [P, P, c_int, c_float, P]),
    'ctgan_mean_bwd': (c_int, [P, P, c_int, c_float, P]),
    'ctgan_softmax_ce_fwd': (c_int, [P, P, P, c_int, c_int, P]),
    'ctgan_softmax_ce_bwd': (c_int, [P, P, P, c_float, P, c_int, c_int, P]),
    'ctgan_adam_step': (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, P, P]),
    'ctgan_peer_alloc': (c_int, [POINTER(P), c_int64]),
    'ctgan_peer_free': (c_int, [P]),
    'ctgan_ipc_get_handle': (c_int, [P, P]),
    'ctgan_ipc_open_handle': (c_int, [P, POINTER(P)]),
    'ctgan_ipc_close_handle': (c_int, [P]),
    'ctgan_peer_flag_bytes': (c_int, []),
    'ctgan_peer_reduce_adam': (c_int, [c_int, c_int, POINTER(P), POINTER(P), POINTER(P), P, P, c_int64, c_float, c_float, c_float, c_float, c_float, P, P]),
    'ctgan_philox_uniform': (c_int, [P, c_int64, c_float, c_float, c_uint64, c_uint64, P, P]),
    'ctgan_counter_add': (c_int, [P, c_uint64, P]),
    'ctgan_philox_normal': (c_int, [P, c_int64, c_uint64, c_uint64, P, P]),
    'ctgan_philox_labels': (c_int, [P, c_int64, c_int, c_uint64, c_uint64, P, P]),
}

EXPORTS = sorted(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)            # AttributeError here == header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


class CtganError(RuntimeError):
    pass


def check(rc, what=''):
    if rc != 0:
        msg = lib.ctgan_last_error()
        raise CtganError('%s failed (rc=%d): %s' % (what or 'libctgan_sm100 call', rc,
                                                   msg.decode('utf-8', 'replace') if msg else ''))


# Count of kernel-launching C-ABI calls made by this process (bench.py reports it).
launch_calls = 0


def call(name, *args):
    global launch_calls
    launch_calls += 1
    check(getattr(lib, name)(*args), name)
