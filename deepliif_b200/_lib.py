"""ctypes binding of libdeepliif_b200.so (the C ABI in include/deepliif_b200.h).  Fails loudly when the
library is missing or does not export a declared symbol — there is no fallback path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdeepliif_b200.so")

FMT_BF16, FMT_FP16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU02, ACT_TANH = 0, 1, 2, 3
PAD_ZERO, PAD_REFLECT = 0, 1


class ConvDesc(C.Structure):
    _fields_ = [("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("nsrc", C.c_int), ("Cin", C.c_int * 2),
                ("Cout", C.c_int), ("R", C.c_int), ("S", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("transposed", C.c_int), ("output_padding", C.c_int), ("pad_mode", C.c_int)]


class FusedSrc(C.Structure):
    """dlb_fused_src: one K-source of dlb_conv_tc_fwd_fused (raw fp32 producer output + norm/act/residual)."""
    _fields_ = [("x", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("act", C.c_int), ("residual", C.c_void_p),
                ("out", C.c_void_p), ("border", C.c_int), ("border_mode", C.c_int)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_cd = C.POINTER(ConvDesc)
_vpp = C.POINTER(C.c_void_p)

# symbol -> (restype, argtypes); mirrors include/deepliif_b200.h one to one
SIGNATURES = {
    "dlb_last_error": (C.c_char_p, []),
    "dlb_version": (_i, []),
    "dlb_release_thread_resources": (_i, []),
    "dlb_conv_out_shape": (_i, [_cd, C.POINTER(_i), C.POINTER(_i)]),
    "dlb_pack_weights_tc": (_i, [_cd, _vp, _i, _vp, _vp, _vp]),
    "dlb_pack_weights_direct": (_i, [_cd, _vp, _vp, _vp]),
    "dlb_conv_tc_fwd": (_i, [_cd, _vpp, _vpp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "dlb_conv_tc_launches": (_i, [_cd, _i, _i, _i]),
    "dlb_conv_tc_fused_mode": (_i, [_cd, _i, _i]),
    "dlb_conv_tc_fwd_fused": (_i, [_cd, C.POINTER(FusedSrc), _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "dlb_conv_tc_fwd_stem": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "dlb_conv_direct_fwd": (_i, [_cd, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "dlb_norm_stats_workspace": (_sz, [_i, _i, _i]),
    "dlb_norm_finalize": (_i, [_vp, _sz, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp]),
    "dlb_norm_stats": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dlb_norm_finalize_bn": (_i, [_vp, _sz, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp]),
    "dlb_norm_stats_bn": (_i, [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _sz, _vp]),
    "dlb_norm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp,
                          _i, _f, C.c_ulonglong, _vp, _vp, _sz, _vp]),
    "dlb_adam_step": (_i, [_vp, _vp, _vp, _vp, C.c_longlong, _f, _f, _f, _f, _i, _f, _vp]),
    "dlb_adam_hyper": (_i, [_f, _f, _f, _i, _f, C.POINTER(_f)]),
    "dlb_adam_step_dev": (_i, [_vp, _vp, _vp, _vp, C.c_longlong, _vp, _f, _f, _f, _vp]),
    "dlb_channel_sum": (_i, [_vp, C.c_longlong, _i, _vp, _i, _vp, _sz, _vp]),
    "dlb_conv_wgrad_workspace": (_sz, [_cd]),
    "dlb_conv_wgrad": (_i, [_cd, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "dlb_head_bwd_pack": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dlb_norm_apply": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, C.c_ulonglong, _vp, _vp]),
    "dlb_stem_window_pack": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dlb_head_finish": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "dlb_stem_conv_weight_bytes": (_sz, []),
    "dlb_stem_conv_pack_weights": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "dlb_stem_conv_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "dlb_head_conv_weight_bytes": (_sz, []),
    "dlb_head_conv_pack_weights": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "dlb_head_conv_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "dlb_tile_luma_sums": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "dlb_u8_to_f32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "dlb_f32_to_u8": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "dlb_seg_finish": (_i, [_vpp, C.POINTER(_f), _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dlb_reflect_fold": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "dlb_stem_window_bwd": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "dlb_cells_posneg_mask": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "dlb_cells_marker_plane": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dlb_cells_mark_background": (_i, [_vp, _i, _i, _vp, _vp]),
    "dlb_cells_label_workspace": (_sz, [_i, _i]),
    "dlb_cells_label": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dlb_cells_stats": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "dlb_cells_classify": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "dlb_cells_enlarge": (_i, [_vp, _vp, _i, _i, _vp]),
    "dlb_cells_final_images": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
}

_lib = None


class DeepliifB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it is not built: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DeepliifB200Error(
            f"{LIB_PATH} not found: build it with `python -m deepliif_b200.build` (nvcc, sm_100a). "
            "deepliif_b200 has no CPU / PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dlb_last_error().decode("utf-8", "replace")
        raise DeepliifB200Error(f"{what} failed (rc={rc}): {msg}")
