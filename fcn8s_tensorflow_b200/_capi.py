"""ctypes binding of libfcn8s_sm100.so (include/fcn8s_b200.h).

Loading never falls back to anything: a missing library raises, and every non-zero status raises Fcn8Error with the
library's own message. PyTorch tensors are used only as device-memory owners (``tensor.data_ptr()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfcn8s_sm100.so")

BF16, F32, BF16X2 = 0, 1, 2
EPI_BIAS, EPI_RELU, EPI_DROPOUT, EPI_MASK, EPI_RESIDUAL, EPI_ROUND_TF32, EPI_COLSUM, EPI_POOL = 1, 2, 4, 8, 16, 64, 128, 256

EXPORTS = [
    "fcn8_version", "fcn8_last_error", "fcn8_device_check", "fcn8_launch_count", "fcn8_debug_set", "fcn8_debug_buffer",
    "fcn8_crc32c", "fcn8_preprocess_im2col",
    "fcn8_conv_gemm_workspace_bytes", "fcn8_conv_gemm", "fcn8_wgrad_gemm_workspace_bytes", "fcn8_wgrad_gemm",
    "fcn8_pack_weights", "fcn8_split_tf32", "fcn8_maxpool_fwd", "fcn8_maxpool_bwd", "fcn8_bias_grad_workspace_bytes",
    "fcn8_bias_grad", "fcn8_head_pack", "fcn8_deconv_cp", "fcn8_deconv_pack", "fcn8_deconv_fwd", "fcn8_deconv_loss",
    "fcn8_deconv_dx", "fcn8_deconv_dw_workspace_bytes", "fcn8_deconv_dw",
    "fcn8_confusion_matrix", "fcn8_adam", "fcn8_l2_reg", "fcn8_shadow_weights",
    "fcn8_set_step_scalars", "fcn8_set_sm_limit", "fcn8_cast_bf16", "fcn8_pack_labels", "fcn8_expand_labels",
    "fcn8_conv1_fwd", "fcn8_conv1_wgrad_workspace_bytes", "fcn8_conv1_wgrad",
]


class Fcn8Error(RuntimeError):
    pass


class PreprocessParams(C.Structure):
    _fields_ = [("images", C.c_void_p), ("out", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("dtype", C.c_int32)]


class ConvParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_lo", C.c_void_p), ("wp", C.c_void_p), ("wp_lo", C.c_void_p),
                ("out", C.c_void_p), ("bias", C.c_void_p), ("mask_src", C.c_void_p), ("residual", C.c_void_p),
                ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
                ("ksize", C.c_int32), ("dtype", C.c_int32), ("nseg", C.c_int32), ("flags", C.c_int32),
                ("mask_scale", C.c_float), ("keep_prob", C.c_float), ("seed", C.c_uint32),
                ("force_splits", C.c_int32), ("force_bn", C.c_int32), ("x_ld", C.c_int32), ("out_ld", C.c_int32),
                ("out_lo", C.c_void_p), ("residual_lo", C.c_void_p), ("w_mode", C.c_int32), ("colsum", C.c_void_p),
                ("seed_ptr", C.c_void_p), ("algo", C.c_int32), ("out_scale", C.c_float), ("colsum_n", C.c_int32),
                ("pool_out", C.c_void_p), ("pool_out_lo", C.c_void_p), ("pool_ld", C.c_int32),
                ("x_sH", C.c_int64), ("x_sN", C.c_int64), ("out_sH", C.c_int64), ("out_sN", C.c_int64)]


class WgradParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_lo", C.c_void_p), ("dy", C.c_void_p), ("dy_lo", C.c_void_p),
                ("dw", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
                ("Cout", C.c_int32), ("ksize", C.c_int32), ("rows_valid", C.c_int32), ("dtype", C.c_int32),
                ("nseg", C.c_int32), ("force_splits", C.c_int32), ("force_bn", C.c_int32), ("x_ld", C.c_int32),
                ("dy_ld", C.c_int32), ("out_cols", C.c_int32), ("out_scale", C.c_float),
                ("x_sH", C.c_int64), ("x_sN", C.c_int64), ("dy_sH", C.c_int64), ("dy_sN", C.c_int64)]


class PackParams(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("out_lo", C.c_void_p), ("ksize", C.c_int32),
                ("Cin", C.c_int32), ("Cout", C.c_int32), ("CinPad", C.c_int32), ("mode", C.c_int32),
                ("dtype", C.c_int32)]


class PoolParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("dx", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("C", C.c_int32), ("dtype", C.c_int32), ("db", C.c_void_p)]


class BiasGradParams(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("db", C.c_void_p), ("P", C.c_int64), ("C", C.c_int32), ("dtype", C.c_int32)]


class Conv1Params(C.Structure):
    _fields_ = [("images", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("w", C.c_void_p),
                ("w_lo", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p), ("out_lo", C.c_void_p),
                ("out_ld", C.c_int32), ("dy", C.c_void_p), ("dy_lo", C.c_void_p), ("dy_ld", C.c_int32),
                ("dw", C.c_void_p), ("pair", C.c_int32)]


class DeconvParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_lo", C.c_void_p), ("x_ld", C.c_int32), ("x_sH", C.c_int64), ("x_sN", C.c_int64),
                ("w", C.c_void_p), ("w_lo", C.c_void_p), ("bias_big", C.c_void_p),
                ("out", C.c_void_p), ("out_lo", C.c_void_p), ("out_ld", C.c_int32), ("out_sH", C.c_int64),
                ("out_sN", C.c_int64), ("skip", C.c_void_p), ("skip_lo", C.c_void_p),
                ("dz", C.c_void_p), ("dz_lo", C.c_void_p), ("dT", C.c_void_p), ("colsum", C.c_void_p),
                ("colsum_n", C.c_int32), ("N", C.c_int32), ("h", C.c_int32), ("wd", C.c_int32), ("C", C.c_int32),
                ("stride", C.c_int32), ("nseg", C.c_int32),
                ("labels", C.c_void_p), ("loss_sum", C.c_void_p), ("dbias", C.c_void_p), ("dz_hi_out", C.c_void_p),
                ("dz_lo_out", C.c_void_p), ("logits", C.c_void_p), ("softmax", C.c_void_p), ("argmax", C.c_void_p),
                ("conf", C.c_void_p), ("grad_scale", C.c_float), ("argmax_u8", C.c_void_p)]


_lib = None


def load():
    """Load the shared library (once). Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Fcn8Error("libfcn8s_sm100.so is not built (run `python -m fcn8s_tensorflow_b200.build`); "
                        "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.fcn8_version.restype = C.c_int32
    lib.fcn8_last_error.restype = C.c_char_p
    lib.fcn8_device_check.argtypes = [C.c_int32]
    lib.fcn8_launch_count.restype = C.c_uint64
    lib.fcn8_debug_set.argtypes = [C.c_int32, C.c_int32]
    lib.fcn8_debug_buffer.argtypes = [C.c_void_p, C.c_int32]
    lib.fcn8_crc32c.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    lib.fcn8_crc32c.restype = C.c_uint32
    lib.fcn8_set_sm_limit.argtypes = [C.c_int32]
    for kv in filter(None, os.environ.get("FCN8_DEBUG", "").split(",")):   # measurement switches, e.g. "3=1"
        k, v = kv.split("=")
        lib.fcn8_debug_set(int(k), int(v))
    vp, sz = C.c_void_p, C.c_size_t
    for name, pt in [("fcn8_conv_gemm", ConvParams), ("fcn8_wgrad_gemm", WgradParams), ("fcn8_bias_grad", BiasGradParams),
                     ("fcn8_deconv_dw", DeconvParams), ("fcn8_conv1_wgrad", Conv1Params)]:
        getattr(lib, name).argtypes = [C.POINTER(pt), vp, sz, vp]
        getattr(lib, name).restype = C.c_int32
        getattr(lib, name + "_workspace_bytes").argtypes = [C.POINTER(pt)]
        getattr(lib, name + "_workspace_bytes").restype = sz
    for name, pt in [("fcn8_preprocess_im2col", PreprocessParams), ("fcn8_pack_weights", PackParams),
                     ("fcn8_maxpool_fwd", PoolParams), ("fcn8_maxpool_bwd", PoolParams),
                     ("fcn8_deconv_fwd", DeconvParams), ("fcn8_deconv_loss", DeconvParams),
                     ("fcn8_deconv_dx", DeconvParams), ("fcn8_conv1_fwd", Conv1Params)]:
        getattr(lib, name).argtypes = [C.POINTER(pt), vp]
        getattr(lib, name).restype = C.c_int32
    i32 = C.c_int32
    lib.fcn8_deconv_cp.argtypes = [i32]
    lib.fcn8_deconv_cp.restype = i32
    lib.fcn8_deconv_pack.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.fcn8_head_pack.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    lib.fcn8_split_tf32.argtypes = [vp, vp, vp, sz, vp]
    lib.fcn8_confusion_matrix.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, vp]
    lib.fcn8_adam.argtypes = [vp, vp, vp, vp, sz, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp,
                              vp]
    lib.fcn8_cast_bf16.argtypes = [vp, vp, sz, vp]
    lib.fcn8_pack_labels.argtypes = [vp, C.c_int64, C.c_int32, vp, C.c_int32]
    lib.fcn8_expand_labels.argtypes = [vp, vp, C.c_int64, C.c_int32, vp]
    lib.fcn8_set_step_scalars.argtypes = [vp, C.c_float, C.c_uint32, vp]
    lib.fcn8_shadow_weights.argtypes = [vp, vp, vp, sz, vp]
    lib.fcn8_l2_reg.argtypes = [vp, vp, vp, sz, C.c_float, vp]
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise Fcn8Error("fcn8 error %d: %s" % (status, load().fcn8_last_error().decode()))


def ptr(t, offset_elems=0):
    """Device pointer of a torch tensor (or None), optionally advanced by a number of elements."""
    return None if t is None else C.c_void_p(t.data_ptr() + offset_elems * t.element_size())
