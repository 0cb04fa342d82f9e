"""ctypes binding of libyad.so (include/yad.h).  There is no fallback: if the library is missing the import of any op fails
loudly with instructions to build it."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libyad.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_RELU, ACT_SIGMOID, ACT_GELU, ACT_HARDSWISH = range(6)
CONV_NORMAL, CONV_TRANSPOSED, CONV_DEFORM = range(3)


class YadTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("ld", C.c_int32)]


class YadEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("img_scale", C.c_void_p), ("pix_scale", C.c_void_p), ("pix_scale_ld", C.c_int32),
                ("act", C.c_int32), ("alpha", C.c_float), ("mul", C.c_void_p), ("mul_ld", C.c_int32), ("add", C.c_void_p),
                ("add_ld", C.c_int32), ("gn_stats", C.c_void_p), ("gn_groups", C.c_int32), ("gate_h", C.c_void_p), ("gate_w", C.c_void_p),
                ("gate_ld", C.c_int32), ("gate_hm", C.c_int32), ("gate_wm", C.c_int32)]


class YadConvDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad_h", C.c_int32),
                ("pad_w", C.c_int32), ("offmask", C.c_void_p), ("offmask_ld", C.c_int32), ("impl", C.c_int32)]


class YadPermuteEntry(C.Structure):
    _fields_ = [("src_off", C.c_int64), ("dst_off", C.c_int64), ("n0", C.c_int64), ("n1", C.c_int64), ("n2", C.c_int64), ("p0", C.c_int64),
                ("p2", C.c_int64), ("s0", C.c_int64), ("s1", C.c_int64), ("s2", C.c_int64), ("flip", C.c_int32), ("dst_f32", C.c_int32)]


class YadMosaicDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_w", C.c_int32), ("x1a", C.c_int32), ("y1a", C.c_int32), ("x2a", C.c_int32), ("y2a", C.c_int32),
                ("x1b", C.c_int32), ("y1b", C.c_int32), ("pad_", C.c_int32)]


class YadImageDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_h", C.c_int32), ("src_w", C.c_int32), ("src_pitch", C.c_int32), ("new_w", C.c_int32),
                ("new_h", C.c_int32), ("top", C.c_int32), ("left", C.c_int32), ("gain", C.c_float), ("pad_x", C.c_float), ("pad_y", C.c_float)]


TP = C.POINTER(YadTensor)
vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/yad.h declares
SIGNATURES = {
    "yad_last_error": (C.c_char_p, []),
    "yad_version": (i32, []),
    "yad_set_pdl": (i32, [i32]),
    "yad_device_is_sm100": (i32, []),
    "yad_struct_size": (i32, [i32]),
    "yad_conv2d": (i32, [TP, vp, C.POINTER(YadConvDesc), C.POINTER(YadEpilogue), TP, i32, vp]),
    "yad_dwconv": (i32, [TP, vp, vp, vp, vp, i32, i32, i32, vp, i32, TP, i32, vp]),
    "yad_gn_stats": (i32, [TP, i32, vp, i32, vp]),
    "yad_gn_apply": (i32, [TP, vp, i32, vp, vp, f32, i32, vp, i32, TP, i32, vp]),
    "yad_sppf_pool": (i32, [TP, TP, TP, TP, i32, vp]),
    "yad_gap": (i32, [TP, vp, i32, vp]),
    "yad_rowcol_mean": (i32, [TP, TP, TP, i32, vp]),
    "yad_rowcol_gate": (i32, [TP, TP, TP, TP, i32, vp]),
    "yad_coordatt_mlp": (i32, [TP, TP, vp, vp, i32, vp, vp, vp, vp, TP, TP, i32, vp]),
    "yad_pool_upsample": (i32, [TP, i32, TP, i32, vp]),
    "yad_mlca_pool": (i32, [TP, vp, i32, i32, vp]),
    "yad_mlca_att": (i32, [vp, vp, vp, i32, f32, i32, i32, i32, vp, vp, vp]),
    "yad_mlca_apply": (i32, [TP, vp, i32, vp, i32, TP, i32, vp]),
    "yad_gate_mlp": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "yad_adt_apply": (i32, [TP, vp, vp, vp, vp, TP, i32, vp]),
    "yad_eltwise": (i32, [i32, TP, vp, i32, vp, i32, vp, i32, f32, f32, f32, TP, i32, vp]),
    "yad_tssa": (i32, [TP, vp, i32, TP, i32, i32, vp]),
    "yad_mha": (i32, [TP, i32, TP, i32, vp]),
    "yad_group_mean": (i32, [TP, i32, TP, i32, vp]),
    "yad_patch_filter": (i32, [TP, vp, f32, vp, i32, TP, i32, vp]),
    "yad_nchw_to_nhwc": (i32, [vp, i32, TP, i32, vp]),
    "yad_u8_to_nhwc": (i32, [vp, i32, TP, f32, i32, vp]),
    "yad_stem_conv": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, i32, TP, i32, vp]),
    "yad_ln_mix": (i32, [TP, vp, vp, vp, vp, f32, TP, i32, vp]),
    "yad_attention_tssa": (i32, [TP, vp, i32, TP, i32, vp]),
    "yad_decode": (i32, [C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                         C.POINTER(f32), i32, i32, i32, i32, vp, vp, i32, vp]),
    "yad_nms_workspace_bytes": (i64, [i32, i32, i32, i32, i32]),
    "yad_nms": (i32, [vp, i32, i32, i32, f32, f32, vp, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp]),
    "yad_letterbox": (i32, [vp, i32, vp, i32, i32, i32, i32, vp]),
    "yad_scale_boxes": (i32, [vp, i32, vp, i32, i32, vp, vp]),
    "yad_val_labels": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "yad_val_match": (i32, [vp, i32, vp, i32, i32, vp, vp, vp, i32, vp, i32, vp, vp, vp, vp]),
    "yad_val_ap_workspace_bytes": (i64, [i64, i32, i32]),
    "yad_val_ap": (i32, [vp, vp, vp, i64, vp, i64, i32, i32, C.c_double, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "yad_tal_workspace_bytes": (i64, [i32, i32, i32]),
    "yad_tal_assign": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "yad_loss_decode": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "yad_loss_bbox": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, f32, f32, vp, vp]),
    "yad_loss_cls": (i32, [vp, vp, i32, i32, i32, vp, f32, vp, vp]),
    "yad_loss_finalize": (i32, [vp, f32, f32, f32, i32, vp, vp]),
    "yad_hsv_lut": (i32, [vp, i32, i32, i32, vp, vp]),
    "yad_flip": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "yad_mosaic4": (i32, [vp, i32, i32, vp, vp]),
    "yad_tc_gemm_selftest": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    # ---- training path
    "yad_eltwise_dev": (i32, [i32, TP, vp, i32, vp, i32, vp, i32, f32, f32, f32, vp, vp, vp, TP, i32, vp]),
    "yad_conv_wgrad": (i32, [TP, TP, C.POINTER(YadConvDesc), vp, i32, vp]),
    "yad_dwconv_wgrad": (i32, [TP, TP, i32, vp, i32, vp]),
    "yad_colsum": (i32, [TP, vp, i32, vp, i32, vp]),
    "yad_dot": (i32, [TP, vp, i32, i32, f32, vp, vp, i32, vp]),
    "yad_dot_pixel": (i32, [TP, vp, i32, TP, i32, vp]),
    "yad_norm_bwd": (i32, [TP, TP, vp, i32, vp, vp, f32, i32, vp, vp, vp, TP, i32, i32, vp]),
    "yad_bn_running_update": (i32, [vp, i32, C.c_double, f32, vp, vp, vp]),
    "yad_act_bwd": (i32, [TP, TP, i32, TP, i32, i32, vp]),
    "yad_bcast_add": (i32, [TP, vp, f32, TP, f32, TP, f32, i32, i32, vp]),
    "yad_rowcol_gate_bwd": (i32, [TP, TP, TP, TP, TP, i32, TP, TP, i32, vp]),
    "yad_mlca_bwd": (i32, [TP, TP, vp, vp, vp, vp, i32, f32, i32, vp, vp, vp, vp, vp, TP, i32, i32, vp]),
    "yad_maxpool5_bwd": (i32, [TP, vp, i32, vp, vp, i32, vp]),
    "yad_cast_acc": (i32, [vp, f32, TP, i32, i32, vp]),
    "yad_pool_upsample_bwd": (i32, [TP, i32, vp, TP, i32, i32, vp]),
    "yad_gate_mlp_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "yad_adt_bwd": (i32, [TP, TP, vp, vp, vp, TP, i32, vp, vp, vp, vp, i32, vp]),
    "yad_gelu_gate_bwd": (i32, [TP, vp, i32, TP, TP, vp, i32, i32, i32, vp]),
    "yad_scale_img": (i32, [TP, vp, TP, i32, i32, vp]),
    "yad_group_mean_bwd": (i32, [TP, i32, TP, i32, i32, vp]),
    "yad_patch_filter_bwd": (i32, [TP, TP, vp, f32, vp, vp, i32, vp]),
    "yad_tssa_bwd": (i32, [TP, vp, i32, TP, i32, TP, vp, i32, vp]),
    "yad_mha_bwd": (i32, [TP, i32, TP, TP, TP, vp, i32, vp]),
    "yad_deform_col": (i32, [TP, TP, TP, i32, vp]),
    "yad_deform_col_bwd": (i32, [TP, TP, TP, vp, TP, i32, vp]),
    "yad_head_pack": (i32, [TP, i32, i32, i32, i32, vp, vp, i32, vp]),
    "yad_head_unpack": (i32, [vp, vp, f32, i32, i32, i32, i32, TP, i32, vp]),
    "yad_fusion_weights": (i32, [vp, i32, vp, vp, vp, vp]),
    "yad_small_gemm": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "yad_permute_pack": (i32, [vp, i32, i64, vp, vp, vp, i32, vp]),
    "yad_permute_unpack": (i32, [vp, i32, i64, vp, vp, vp]),
    "yad_sqnorm": (i32, [vp, i64, vp, vp]),
    "yad_sgd_step": (i32, [vp, vp, vp, vp, i64, C.POINTER(f32), C.POINTER(f32), f32, f32, vp, i32, vp]),
    "yad_adamw_step": (i32, [vp, vp, vp, vp, vp, i64, C.POINTER(f32), C.POINTER(f32), f32, f32, f32, i32, f32, vp, vp]),
    "yad_ema_update": (i32, [vp, vp, i64, f32, vp]),
    "yad_sgd_step_dev": (i32, [vp, vp, vp, vp, i64, vp, vp, vp]),
    "yad_adamw_step_dev": (i32, [vp, vp, vp, vp, vp, i64, vp, vp, vp]),
    "yad_ema_update_dev": (i32, [vp, vp, i64, vp, vp]),
}

_lib = None


def load():
    """Load libyad.so and bind every symbol.  Raises (never falls back) when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m yolo_ad_refine_b200.build` "
                           "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the YOLO-AD-Refine hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class YadError(RuntimeError):
    pass


def check(status, what=""):
    if status != 0:
        raise YadError(f"{what}: {load().yad_last_error().decode(errors='replace')}")
