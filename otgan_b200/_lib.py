"""ctypes binding of libotgan.so (include/otgan.h).  There is NO fallback: if the CUDA library is missing or a call
fails, this module raises -- the product path never routes through a CPU/oracle implementation."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libotgan.so")

OTGAN_MAX_BLOCKS = 8
OTGAN_MAX_TERMS = 3
OTGAN_MAX_OUTPUTS = 8
COST_COSINE, COST_EUCLID_MEAN = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


class Plan(ctypes.Structure):
    """otgan_plan_t"""
    _fields_ = [("n_out", _i),
                ("nterms", _i * OTGAN_MAX_OUTPUTS),
                ("blk", (_i * OTGAN_MAX_TERMS) * OTGAN_MAX_OUTPUTS),
                ("trans", (_i * OTGAN_MAX_TERMS) * OTGAN_MAX_OUTPUTS),
                ("src", (_i * OTGAN_MAX_TERMS) * OTGAN_MAX_OUTPUTS),
                ("coef", (_f * OTGAN_MAX_TERMS) * OTGAN_MAX_OUTPUTS)]


class DenseGeom(ctypes.Structure):
    """otgan_dense_geom_t"""
    _fields_ = [("B", _i), ("H", _i), ("W", _i), ("n_base", _i), ("base_ch", _i * 4), ("L", _i), ("growth", _i)]


_ll = ctypes.c_longlong
_gp = ctypes.POINTER(DenseGeom)

# name -> (restype, argtypes); must list every symbol include/otgan.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "otgan_abi_version": (_i, []),
    "otgan_last_error": (ctypes.c_char_p, []),
    "otgan_launch_count": (ctypes.c_uint64, []),
    "otgan_reset_launch_count": (None, []),
    "otgan_workspace_bytes_cost": (_sz, [_i, _i, _i, _i, _i]),
    "otgan_cost_blocks_f32": (_i, [_i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _f, _vp, _vp, _sz, _i, _vp]),
    "otgan_sinkhorn_f32": (_i, [_i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "otgan_sinkhorn_ex_f32": (_i, [_i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "otgan_workspace_bytes_plan": (_sz, []),
    "otgan_workspace_bytes_plan_h": (_sz, [_i]),
    "otgan_plan_apply_f32": (_i, [ctypes.POINTER(Plan), _i, _i, _vp, _vp, _i, _vp, _i, _vp, _sz, _i, _vp]),
    "otgan_matched_two_batch_f32": (_i, [_i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _i, _vp]),
    "otgan_grad_features_f32": (_i, [_i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _sz, _i, _vp]),
    "otgan_grad_features_rows_f32": (_i, [_i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _sz, _i, _vp]),
    "otgan_matched_single_batch_f32": (_i, [_i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _i, _vp]),
    "otgan_workspace_bytes_distance": (_sz, [_i, _i]),
    "otgan_calc_distance_f32": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _sz, _vp]),
    "otgan_distance_from_pc_f32": (_i, [_vp, _vp, _i, _vp, _vp]),
    "otgan_adam_ema_f32": (_i, [_sz, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _vp]),
    "otgan_adam_ema_dev_f32": (_i, [_sz, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp]),
    "otgan_crelu_l2norm_fwd_f32": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "otgan_crelu_l2norm_bwd_f32": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "otgan_workspace_bytes_weightnorm": (_sz, [_i, _i]),
    "otgan_weightnorm_fwd_f32": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_weightnorm_bwd_f32": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_crelu_pad_fwd_f32": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "otgan_crelu_pad_bwd_f32": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "otgan_glu_up_fwd_f32": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "otgan_glu_up_bwd_f32": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "otgan_workspace_bytes_conv_gemm": (_sz, [_i, _i, _i, _i]),
    "otgan_conv2d_fprop_tf32": (_i, [_i] * 10 + [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_conv2d_fprop_crelu_tf32": (_i, [_i] * 10 + [_vp, _vp, _vp, _vp, _vp]),
    "otgan_crelu_bwd_from_activated_f32": (_i, [ctypes.c_longlong, _i, _vp, _vp, _vp, _vp]),
    "otgan_conv2d_dgrad_tf32": (_i, [_i] * 10 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_workspace_bytes_conv_wgrad": (_sz, [_i] * 8),
    "otgan_conv2d_wgrad_tf32": (_i, [_i] * 10 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_conv_set_option": (_i, [_i, _i]),
    "otgan_conv_plan_describe": (_i, [_i] * 11 + [ctypes.POINTER(ctypes.c_longlong), _i]),
    "otgan_ohwi_to_ihwo_f32": (_i, [_i, _i, _i, _vp, _vp, _vp]),
    "otgan_up2_subtaps": (_i, [_i, _i]),
    "otgan_up2_weight_presum_f32": (_i, [_i] * 6 + [_vp, _vp, _vp]),
    "otgan_up2_weight_unsum_f32": (_i, [_i] * 6 + [_vp, _vp, _vp]),
    "otgan_conv2d_up2_fprop_tf32": (_i, [_i] * 9 + [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_conv2d_up2_dgrad_tf32": (_i, [_i] * 9 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_workspace_bytes_conv_up2_wgrad": (_sz, [_i] * 9),
    "otgan_conv2d_up2_wgrad_tf32": (_i, [_i] * 9 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_im2col_narrow_f32": (_i, [_i] * 9 + [_vp, _vp, _i, _vp]),
    "otgan_col2im_narrow_f32": (_i, [_i] * 9 + [_vp, _i, _vp, _vp, _vp]),
    "otgan_workspace_bytes_colsum": (_sz, [_i, _i]),
    "otgan_colsum_f32": (_i, [_i, _i, _vp, _vp, _vp, _sz, _vp]),
    "otgan_weightnorm_fwd2_f32": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_weightnorm_bwd_hwio_f32": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_conv2d_wgrad_hwio_tf32": (_i, [_i] * 10 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_conv2d_up2_wgrad_hwio_tf32": (_i, [_i] * 9 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_up2_weight_presum_ihwo_f32": (_i, [_i] * 6 + [_vp, _vp, _vp]),
    "otgan_up2_weight_unsum_hwio_f32": (_i, [_i] * 6 + [_vp, _vp, _vp]),
    "otgan_conv2d_fprop_ex_tf32": (_i, [_i] * 12 + [_vp, _vp, _vp, _vp, _i, _vp]),
    "otgan_conv2d_dgrad_ex_tf32": (_i, [_i] * 12 + [_vp, _vp, _vp, _vp]),
    "otgan_workspace_bytes_conv_wgrad_ex": (_sz, [_i] * 8),
    "otgan_conv2d_wgrad_ex_tf32": (_i, [_i] * 12 + [_vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_crelu8_fwd_f32": (_i, [_ll, _i, _vp, _i, _vp, _i, _vp]),
    "otgan_crelu8_bwd_f32": (_i, [_ll, _i, _vp, _i, _vp, _i, _vp, _i, _vp]),
    "otgan_crelu8_perm_host": (_i, [_i, ctypes.POINTER(_i), _i, ctypes.POINTER(_i), _i]),
    "otgan_weightnorm_fwd_ex_f32": (_i, [_i, _i, _vp, _vp, _vp, _i, _ll, _ll, _vp, _vp, _vp, _sz, _vp]),
    "otgan_weightnorm_bwd_ex_f32": (_i, [_i, _i, _vp, _vp, _vp, _vp, _i, _ll, _ll, _vp, _vp, _vp, _vp, _sz, _vp]),
    "otgan_dense_channels": (_i, [_gp]),
    "otgan_dense_wb_floats": (_sz, [_gp]),
    "otgan_dense_build_wb_f32": (_i, [_gp, _vp, _vp, _vp]),
    "otgan_dense_block_fprop_tf32": (_i, [_gp, _vp, _vp, _vp, _vp, _vp]),
    "otgan_workspace_bytes_dense_bgrad": (_sz, [_gp]),
    "otgan_dense_block_bgrad_tf32": (_i, [_gp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


class OtganError(RuntimeError):
    pass


def load():
    """Load libotgan.so (built in-tree by `python -m otgan_b200.build`).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "otgan_b200: %s not found -- the CUDA extension is required (there is no CPU fallback). "
            "Build it with `python -m otgan_b200.build` (needs nvcc, targets sm_100a)." % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().otgan_last_error()
        raise OtganError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(int(p)) for p in ptrs])


def float_array(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def launch_count():
    return int(load().otgan_launch_count())


def reset_launch_count():
    load().otgan_reset_launch_count()
