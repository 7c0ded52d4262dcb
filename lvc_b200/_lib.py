"""ctypes binding of liblvcb200.so (the C ABI declared in include/lvcb200.h).

The CUDA library is the product: there is NO fallback.  Importing this module without the built library, or
calling an op with CPU tensors, raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LVCB200_LIB") or os.path.join(_HERE, "liblvcb200.so")   # LVCB200_LIB: A/B another build on the same box

ABI_VERSION = 2
F32, BF16, U8, F16 = 0, 1, 2, 3
OUT_NCHW, OUT_NHWC = 0, 1


class FMap(Structure):
    _fields_ = [("base", c_void_p), ("H", c_int), ("W", c_int), ("img_stride", c_int64), ("row_stride", c_int64),
                ("c_stride", c_int64), ("spatial_scale", c_float)]


class RpnLevel(Structure):
    _fields_ = [("logits", c_void_p), ("deltas", c_void_p), ("H", c_int), ("W", c_int), ("A", c_int), ("stride", c_int),
                ("img_stride_l", c_int64), ("row_stride_l", c_int64), ("pix_stride_l", c_int64),
                ("img_stride_d", c_int64), ("row_stride_d", c_int64), ("pix_stride_d", c_int64),
                ("cell_anchors", c_float * 12)]


class RpnParams(Structure):
    _fields_ = [("n_images", c_int), ("n_levels", c_int), ("pre_nms_topk", c_int), ("post_nms_topk", c_int),
                ("nms_thresh", c_float), ("min_box_size", c_float), ("weights", c_float * 4), ("nms_mode", c_int)]


class DetParams(Structure):
    _fields_ = [("n_images", c_int), ("num_classes", c_int), ("max_rois_per_image", c_int), ("class_agnostic", c_int),
                ("weights", c_float * 4), ("score_thresh", c_float), ("nms_thresh", c_float), ("topk_per_image", c_int),
                ("nms_mode", c_int)]


class GemmDesc(Structure):
    _fields_ = [("a_dtype", c_int), ("A", c_void_p), ("lda", c_int64), ("M_rows", c_int64), ("W", c_void_p), ("ldw", c_int64),
                ("bias", c_void_p), ("residual", c_void_p), ("ldr", c_int64), ("D", c_void_p), ("ldd", c_int64),
                ("d_dtype", c_int), ("M", c_int64), ("N", c_int), ("K", c_int), ("taps", c_int), ("shift", c_int32 * 9),
                ("relu", c_int), ("plane_h", c_int), ("plane_w", c_int),
                ("upsample_add", c_void_p), ("ldu", c_int64), ("up_plane_h", c_int), ("up_plane_w", c_int),
                ("split_rows", c_int64)]


class ChainPlan(Structure):
    _fields_ = [("workspace", c_void_p), ("counter_bytes", c_int64), ("total_tiles", c_int64), ("n_layers", c_int), ("grid", c_int)]


_SIGS = {
    "lvcb200_abi_version": (c_int, []),
    "lvcb200_last_error": (c_char_p, []),
    "lvcb200_launch_count": (c_int64, []),
    "lvcb200_roi_align_nchw_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_float,
                                           c_int, c_int, c_void_p, c_void_p]),
    "lvcb200_assign_boxes_to_levels": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lvcb200_roi_pool_fpn": (c_int, [POINTER(FMap), c_int, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                     c_int, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "lvcb200_batched_nms_workspace": (c_size_t, [c_int64]),
    "lvcb200_batched_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "lvcb200_rpn_proposals_workspace": (c_size_t, [POINTER(RpnParams)]),
    "lvcb200_rpn_proposals": (c_int, [POINTER(RpnLevel), POINTER(RpnParams), c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "lvcb200_detections_workspace": (c_size_t, [POINTER(DetParams)]),
    "lvcb200_detections": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                   POINTER(DetParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "lvcb200_apply_deltas_clip": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, POINTER(c_float), c_int, c_void_p,
                                          c_void_p]),
    "lvcb200_knn_prepared_bytes": (c_size_t, [c_int, c_int]),
    "lvcb200_knn_prepare": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "lvcb200_knn_verify": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "lvcb200_knn_prepare_euclid": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "lvcb200_knn_verify_euclid": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    "lvcb200_knn_tc_select": (c_int, [c_int]),
    "lvcb200_knn_tc_workspace": (c_size_t, [c_int64, c_int, c_int]),
    "lvcb200_knn_verify_tc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "lvcb200_gemm_bf16": (c_int, [POINTER(GemmDesc), c_void_p]),
    "lvcb200_gemm_chain_workspace": (c_size_t, [POINTER(GemmDesc), c_int]),
    "lvcb200_gemm_chain_plan": (c_int, [POINTER(GemmDesc), c_int, c_void_p, c_size_t, POINTER(ChainPlan)]),
    "lvcb200_gemm_chain_run": (c_int, [POINTER(ChainPlan), c_void_p]),
    "lvcb200_roi_align_backward_nchw_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                                    c_int, c_void_p, c_void_p]),
    "lvcb200_pairwise_iou": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "lvcb200_match_boxes_workspace": (c_size_t, [c_int64]),
    "lvcb200_match_boxes": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, POINTER(c_float), c_int, POINTER(ctypes.c_int8), c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "lvcb200_gather_rows_p2p": (c_int, [POINTER(c_void_p), c_int, POINTER(ctypes.c_int32), c_int, c_void_p, c_void_p]),
    "lvcb200_subsample_labels": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, ctypes.c_double, c_int64, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "lvcb200_rpn_losses": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, POINTER(c_float), c_float, c_void_p,
                                   c_void_p]),
    "lvcb200_fast_rcnn_losses": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, POINTER(c_float), c_float,
                                         c_void_p, c_void_p]),
    "lvcb200_stem_s2d4": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lvcb200_maxpool_s2d": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lvcb200_crops_qe": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lvcb200_subsample2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lvcb200_upsample2_add": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "lvcb200_stem_s2d4_pair": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "lvcb200_maxpool_s2d_pair": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "lvcb200_upsample2_add_pair": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "lvcb200_pair_merge": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "lvcb200_pair_split": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "lvcb200_row_inv_norm": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "lvcb200_vit_patchify": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lvcb200_vit_assemble": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lvcb200_layernorm": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int64, c_void_p]),
    "lvcb200_gelu": (c_int, [c_void_p, c_int64, c_void_p]),
    "lvcb200_attention": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "lvcb200_attention_tc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "lvcb200_candidate_filter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "lvcb200_make_rois": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}
EXPORTS = tuple(_SIGS)

_lib = None


class LvcB200Error(RuntimeError):
    pass


def load():
    """dlopen liblvcb200.so; raises if it has not been built (python -m lvc_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LvcB200Error(f"{LIB_PATH} is missing: build it with `python -m lvc_b200.build` "
                               "(there is no CPU or PyTorch fallback for these ops)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.lvcb200_abi_version() != ABI_VERSION:
            raise LvcB200Error("liblvcb200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().lvcb200_last_error().decode(errors="replace")
        raise LvcB200Error(f"{what} failed (code {rc}): {msg}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise LvcB200Error("lvc_b200 ops run on CUDA tensors only (no CPU fallback); got a CPU tensor")


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def launch_count():
    return int(load().lvcb200_launch_count())
