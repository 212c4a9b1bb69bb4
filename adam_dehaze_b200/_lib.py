"""ctypes binding of libadb200.so — the C-ABI declared in include/adb200.h.

There is no CPU fallback: if the library is missing it is built (nvcc) and if that fails, or a call returns a
non-zero status, an exception is raised.  Nothing here imports the oracle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libadb200.so")

# --- enums (mirror include/adb200.h)
ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_SIGMOID2 = 0, 1, 2, 3, 4
CONV_S1, CONV_S2, CONVT_4X4S2, CONV_K4_S2D = 0, 1, 2, 3
EPI_FEATURE, EPI_DOT, EPI_IMAGE = 0, 1, 2
IMG_BLEND, IMG_RESIDUAL, IMG_GUIDED = 0, 1, 2


class AdbError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("src0", C.c_void_p), ("c0", C.c_int32), ("c0_pitch", C.c_int32),
        ("src1", C.c_void_p), ("c1", C.c_int32), ("c1_pitch", C.c_int32),
        ("n", C.c_int32), ("h_in", C.c_int32), ("w_in", C.c_int32),
        ("kind", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("pad", C.c_int32),
        ("w_packed", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("cout", C.c_int32), ("cout_pad", C.c_int32),
        ("act", C.c_int32),
        ("epi", C.c_int32),
        ("residual", C.c_void_p), ("res_pitch", C.c_int32),
        ("dst", C.c_void_p), ("dst_pitch", C.c_int32), ("dst_c_off", C.c_int32),
        ("dot_w", C.c_void_p), ("dot_b", C.c_float), ("dot_out", C.c_void_p),
        ("img_mode", C.c_int32),
        ("img_x", C.c_void_p), ("img_out", C.c_void_p), ("img_index", C.c_void_p),
        ("img_guidance", C.c_void_p), ("img_alpha", C.c_void_p),
        ("n_dev", C.c_void_p), ("n_start", C.c_int32),
        ("tune_mt", C.c_int32), ("tune_stages", C.c_int32), ("tune_acc_stages", C.c_int32),
        ("tune_flags", C.c_int32),
        ("pre_scale", C.c_void_p), ("pre_shift", C.c_void_p),
        ("w_fold", C.c_void_p),
        ("stat_out", C.c_void_p), ("stat_mode", C.c_int32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("grad", C.c_void_p), ("cg", C.c_int32), ("cg_pitch", C.c_int32), ("cg_true", C.c_int32),
        ("act0", C.c_void_p), ("c0", C.c_int32), ("c0_pitch", C.c_int32),
        ("act1", C.c_void_p), ("c1", C.c_int32), ("c1_pitch", C.c_int32),
        ("n", C.c_int32), ("h_in", C.c_int32), ("w_in", C.c_int32),
        ("kind", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("pad", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("dw", C.c_void_p), ("layout", C.c_int32), ("stem_kw", C.c_int32), ("accumulate", C.c_int32),
        ("mode", C.c_int32),
    ]


class F32ConvDesc(C.Structure):
    _fields_ = [
        ("flag_index", C.c_void_p), ("flag_count", C.c_void_p), ("cursor", C.c_void_p), ("cap", C.c_int32),
        ("x", C.c_void_p), ("x_slot", C.c_void_p), ("in_nchw", C.c_int32),
        ("h_in", C.c_int32), ("w_in", C.c_int32), ("cin", C.c_int32), ("in_pitch", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("w", C.c_void_p), ("cout", C.c_int32),
        ("pre_scale", C.c_void_p), ("pre_shift", C.c_void_p), ("post_scale", C.c_void_p), ("post_shift", C.c_void_p),
        ("post_relu", C.c_int32),
        ("residual", C.c_void_p), ("res_pitch", C.c_int32),
        ("y", C.c_void_p), ("out_pitch", C.c_int32), ("out_c_off", C.c_int32),
    ]


WG_OIHW, WG_STEM = 0, 1

_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> argtypes (all return int status unless listed in _SPECIAL)
SIGNATURES = {
    "adb_version": [],
    "adb_device_check": [],
    "adb_kernel_error_flag": [],
    "adb_conv2d": [C.POINTER(ConvDesc), _P],
    "adb_debug_timeline": [_P, _I],
    "adb_wgrad": [C.POINTER(WgradDesc), _P],
    "adb_bn_finalize_stats": [_P, _L, _I, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "adb_bn_train_stats": [_P, _L, _I, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "adb_affine_act": [_P, _I, _L, _I, _P, _P, _P, _I, _I, _P, _I, _P],
    "adb_bn_bwd": [_P, _I, _P, _I, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _I, _P],
    "adb_bn_relu_bwd": [_P, _I, _P, _I, _L, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P],
    "adb_add_bf16": [_P, _I, _P, _I, _L, _I, _P],
    "adb_img_head_fwd": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "adb_img_head_bwd": [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "adb_dot_head_fwd": [_P, _I, _I, _P, _P, _L, _P, _P],
    "adb_dot_head_bwd": [_P, _P, _P, _I, _I, _P, _L, _P, _I, _P, _P],
    "adb_attn_bwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "adb_blend3_bwd": [_P, _P, _P, _P, _P, _F, _I, _L, _P, _P, _P, _P, _P, _P],
    "adb_image_affine": [_P, _I, _I, _I, _P, _P, _P, _P],
    "adb_maxpool_fwd": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adb_maxpool_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adb_mse_feat": [_P, _P, _L, _F, _P, _P, _P],
    "adb_lpips_tap": [_P, _P, _I, _I, _I, _I, _P, _F, _P, _P, _P],
    "adb_stem_unpack": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "adb_broadcast_hw": [_P, _I, _I, _I, _I, _F, _P, _P],
    "adb_mul_f32": [_P, _P, _P, _L, _P, _P],
    "adb_linear_bwd": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "adb_image_metrics": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "adb_avgpool2x2_bwd": [_P, _I, _I, _I, _I, _I, _P, _I, _P],
    "adb_gather_cast": [_P, _P, _L, _P, _P],
    "adb_gather_cast_multi": [_P, _I, _L, _P],
    "adb_upsample_bilinear_bwd": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adb_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P],
    "adb_adam_step_segments": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _P, _I, _P, _P, _P, _P, _P],
    "adb_stem_pack": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "adb_nchw_to_nhwc_bf16": [_P, _I, _I, _I, _I, _I, _P, _P],
    "adb_nhwc_bf16_to_nchw": [_P, _I, _I, _I, _I, _I, _P, _P],
    "adb_attn_pool_from_stats": [_P, _I, _I, _I, _I, _P, _I, _P, _P, _P],
    "adb_attn_pool": [_P, _I, _I, _I, _I, _P, _I, _P, _P],
    "adb_attn_gate_stats": [_P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _I, _P, _P, _P],
    "adb_attn_apply": [_P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P],
    "adb_maxpool3x3s2": [_P, _I, _I, _I, _I, _P, _I, _P],
    "adb_global_avgpool": [_P, _I, _I, _I, _I, _P, _P, _P],
    "adb_affine_relu": [_P, _L, _I, _I, _P, _P, _P, _I, _P],
    "adb_avgpool2x2": [_P, _I, _I, _I, _I, _I, _P, _I, _P],
    "adb_maxpool_kxk": [_P, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "adb_upsample_bilinear": [_P, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P],
    "adb_head_mlp": [_P, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P],
    "adb_linear": [_P, _I, _I, _P, _P, _I, _I, _P, _P],
    "adb_route": [_P, _P, _I, _I, _P, _P, _P, _P, _P],
    "adb_guard_flags": [_P, _I, _I, _F, _P, _P, _P, _P],
    "adb_guard_set_slots": [_P, _P, _P, _P],
    "adb_f32_conv2d": [C.POINTER(F32ConvDesc), _P],
    "adb_f32_pool": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _P],
    "adb_f32_global_avgpool": [_P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P],
    "adb_guard_scatter": [_P, _P, _P, _I, _P, _I, _P, _P],
    "adb_guard_advance": [_P, _P, _I, _P],
    "adb_guard_graph_begin": [_P, _P, _I, _P, C.POINTER(C.c_void_p)],
    "adb_guard_graph_end": [_P],
    "adb_guard_graph_launch": [_P, _P],
    "adb_guard_graph_destroy": [_P],
    "adb_image_u8_to_f32": [_P, _I, _I, _I, _L, _I, _P, _P, _I, _I, _P],
    "adb_zero_unrouted": [_P, _P, _I, _L, _P],
    "adb_blend3": [_P, _P, _P, _P, _F, _I, _L, _P, _P, _P],
    "adb_l1_mse_fwd": [_P, _P, _L, _P, _P],
    "adb_l1_bwd": [_P, _P, _L, _F, _P, _P],
    "adb_mse_bwd": [_P, _P, _L, _F, _P, _P],
    "adb_ce_fwd_bwd": [_P, _P, _I, _I, _F, _P, _P, _P],
}
_SPECIAL = {
    "adb_last_error": ([], C.c_char_p),
    "adb_conv2d_flops": ([C.POINTER(ConvDesc)], C.c_double),
    "adb_conv2d_stat_slots": ([C.POINTER(ConvDesc)], C.c_int64),
    "adb_pool_scratch_floats": ([_I, _I, _I, _I], C.c_int64),
    "adb_attn_pool_stat_scratch_floats": ([_I, _I, _I], C.c_int64),
    "adb_wgrad_workspace_bytes": ([C.POINTER(WgradDesc)], C.c_int64),
    "adb_bn_scratch_floats": ([_L, _I], C.c_int64),
    "adb_bn_stat_scratch_floats": ([_L, _I], C.c_int64),
    "adb_attn_bwd_scratch_floats": ([_I, _I, _I, _I], C.c_int64),
    "adb_wgrad_flops": ([C.POINTER(WgradDesc)], C.c_double),
}
EXPORTED_SYMBOLS = sorted(list(SIGNATURES) + list(_SPECIAL))

_lib = None


def lib_path():
    return _LIB_PATH


def load(build_if_missing=True):
    """Load (building first if needed) and return the ctypes library object."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH) and not build_if_missing:
        raise AdbError(f"{_LIB_PATH} is missing; run `python -m adam_dehaze_b200.build`")
    if build_if_missing:
        # build() is a no-op when the library's stamp matches the digest of csrc/ + include/ + flags; after an edit of the
        # sources it rebuilds instead of silently loading the stale library
        from . import build as _build
        _build.build()
    lib = C.CDLL(_LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error():
    msg = load().adb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status, what=""):
    if status != 0:
        raise AdbError(f"{what or 'libadb200'} failed with status {status}: {last_error()}")


def call(name, *args):
    """Call an int-status entry point and raise AdbError on failure."""
    fn = getattr(load(), name)
    check(fn(*args), name)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    """The calling thread's current CUDA stream as a raw handle.  Read through torch's C accessor: `torch.cuda.current_stream()`
    builds a Stream object through several Python layers (~15 us, x ~1.4 k calls per training step)."""
    import torch
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
