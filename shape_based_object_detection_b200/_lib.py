"""ctypes binding of libsbod.so (the C ABI declared in include/sbod.h).

There is no CPU fallback: if the shared library is missing, or a tensor is not on a CUDA device,
the call raises. PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsbod.so")

# ---- enums (mirror include/sbod.h) ----------------------------------------------------------
IOU_METRICS, IOU_JACCARD, IOU_INTERSECT = 0, 1, 2
BOX_XY_TO_CXCY, BOX_CXCY_TO_XY = 0, 1
CODEC_TRANSFORMS, CODEC_IOU_UTILS = 0, 1
PAIR_IOU, PAIR_GIOU, PAIR_DIOU, PAIR_CIOU = 0, 1, 2, 3
REG_L1_ELEM_MEAN, REG_SMOOTH_L1, REG_IOU, REG_GIOU, REG_DIOU, REG_CIOU = 0, 1, 2, 3, 4, 5
CLS_CE_MINE_NONPOS, CLS_CE_MINE_NEG, CLS_CE_MINE_BATCH, CLS_FOCAL_SUM, CLS_FOCAL_NORM = 0, 1, 2, 3, 4
ACT_SOFTMAX, ACT_SIGMOID, ACT_NONE = 0, 1, 2
BOX_OFFSET, BOX_CENTER, BOX_CORNER = 0, 1, 2

EXPORTS = [
    "sbod_abi_version", "sbod_error_string", "sbod_iou_matrix", "sbod_box_convert",
    "sbod_box_encode", "sbod_box_decode", "sbod_offset2bbox", "sbod_arm_easy_negative", "sbod_pair_iou_fwd",
    "sbod_pair_iou_bwd", "sbod_smooth_l1", "sbod_softmax_focal", "sbod_sigmoid_focal",
    "sbod_nms_workspace_bytes", "sbod_nms", "sbod_assign_workspace_bytes", "sbod_assign",
    "sbod_match_workspace_bytes", "sbod_match", "sbod_loss_workspace_bytes", "sbod_workspace_init",
    "sbod_loss_forward", "sbod_loss_forward_stage", "sbod_detect_stage", "sbod_loss_finalize", "sbod_loss_backward", "sbod_loss_targets",
    "sbod_detect_workspace_bytes", "sbod_detect_workspace_zero_bytes", "sbod_detect",
    "sbod_loss_forward_host_arena_bytes", "sbod_loss_forward_host",
    "sbod_fcos_workspace_bytes", "sbod_fcos_forward", "sbod_fcos_backward", "sbod_fcos_postprocess",
    "sbod_selftest_div", "sbod_map_workspace_bytes", "sbod_map", "sbod_bce_focal", "sbod_diou_nms",
]


class LossDesc(C.Structure):
    _fields_ = [
        ("locs", C.c_void_p), ("scores", C.c_void_p), ("priors_cxcy", C.c_void_p),
        ("priors_xy", C.c_void_p), ("anchors_xy", C.c_void_p), ("gt_boxes", C.c_void_p),
        ("gt_labels", C.c_void_p), ("gt_offsets", C.c_void_p), ("exclude", C.c_void_p),
        ("N", C.c_int32), ("P", C.c_int32), ("C", C.c_int32), ("gmax", C.c_int32),
        ("thr_pos", C.c_float), ("thr_neg", C.c_float),
        ("reg_kind", C.c_int32), ("cls_kind", C.c_int32), ("binarize_labels", C.c_int32),
        ("neg_pos_ratio", C.c_int32), ("reg_weight", C.c_float), ("smooth_l1_beta", C.c_float),
        ("focal_alpha", C.c_float), ("focal_gamma", C.c_float),
        ("ov", C.c_void_p), ("obj", C.c_void_p), ("lse", C.c_void_p), ("ce", C.c_void_p),
        ("sel", C.c_void_p), ("partials", C.c_void_p), ("sums", C.c_void_p), ("loss", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("grad_scores_prefill", C.c_void_p),
    ]


class DetectDesc(C.Structure):
    _fields_ = [
        ("locs", C.c_void_p), ("scores", C.c_void_p), ("priors_cxcy", C.c_void_p),
        ("prior_keep", C.c_void_p),
        ("N", C.c_int32), ("P", C.c_int32), ("C", C.c_int32),
        ("act_kind", C.c_int32), ("box_kind", C.c_int32), ("clamp_inplace", C.c_int32),
        ("min_score", C.c_float), ("max_overlap", C.c_float), ("top_k", C.c_int32),
        ("second_nms_thr", C.c_float), ("pre_nms_topk", C.c_int32),
        ("out_boxes", C.c_void_p), ("out_labels", C.c_void_p), ("out_scores", C.c_void_p),
        ("out_prior", C.c_void_p), ("out_counts", C.c_void_p), ("out_cap", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class FcosDesc(C.Structure):
    _fields_ = [
        ("locs", C.c_void_p), ("scores", C.c_void_p), ("centerness", C.c_void_p), ("locations", C.c_void_p),
        ("loc_aux", C.c_void_p), ("gt_boxes", C.c_void_p), ("gt_labels", C.c_void_p), ("gt_offsets", C.c_void_p),
        ("N", C.c_int32), ("P", C.c_int32), ("C", C.c_int32), ("center_sample", C.c_int32),
        ("reg_weight", C.c_float), ("focal_alpha", C.c_float), ("focal_gamma", C.c_float),
        ("lab", C.c_void_p), ("tgt", C.c_void_p), ("sums", C.c_void_p), ("loss", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class SbodError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def _declare(lib):
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.sbod_abi_version.restype = C.c_int
    lib.sbod_error_string.restype = C.c_char_p
    lib.sbod_error_string.argtypes = [C.c_int]
    sigs = {
        "sbod_iou_matrix": [vp, i32, vp, i32, i32, vp, vp],
        "sbod_box_convert": [vp, vp, i32, i32, vp],
        "sbod_box_encode": [vp, vp, vp, i32, i32, f32, f32, vp],
        "sbod_box_decode": [vp, vp, vp, i32, i32, f32, f32, vp],
        "sbod_offset2bbox": [vp, vp, vp, vp, i32, i32, vp],
        "sbod_arm_easy_negative": [vp, C.c_longlong, f32, vp, vp],
        "sbod_selftest_div": [vp, vp, C.c_longlong, vp, vp, vp, vp],
        "sbod_map": [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                     C.POINTER(C.c_float), vp, vp, C.c_size_t, vp],
        "sbod_pair_iou_fwd": [vp, vp, i32, i32, vp, vp],
        "sbod_pair_iou_bwd": [vp, vp, vp, i32, i32, vp, vp, vp],
        "sbod_smooth_l1": [vp, vp, i32, f32, vp, vp, vp],
        "sbod_softmax_focal": [vp, vp, i32, i32, f32, f32, f32, vp, vp, vp],
        "sbod_sigmoid_focal": [vp, vp, i32, i32, f32, f32, vp, vp, vp],
        "sbod_bce_focal": [vp, vp, i32, i32, f32, f32, vp, vp, vp],
        "sbod_nms": [vp, vp, i32, f32, i32, vp, vp, vp, sz, vp],
        "sbod_diou_nms": [vp, vp, i32, f32, i32, f32, vp, vp, vp, sz, vp],
        "sbod_assign": [vp, vp, vp, i32, i32, vp, i32, i32, f32, f32, vp, vp, vp, vp, vp, sz, vp],
        "sbod_match": [f32, vp, i32, vp, i32, f32, f32, vp, i32, vp, vp, vp, sz, vp],
        "sbod_workspace_init": [vp, sz, vp],
        "sbod_loss_forward": [C.POINTER(LossDesc), vp],
        "sbod_loss_finalize": [C.POINTER(LossDesc), vp],
        "sbod_loss_forward_stage": [C.POINTER(LossDesc), i32, vp],
        "sbod_detect_stage": [C.POINTER(DetectDesc), i32, vp],
        "sbod_loss_backward": [C.POINTER(LossDesc), vp, vp, vp, vp],
        "sbod_loss_targets": [C.POINTER(LossDesc), vp, vp, vp],
        "sbod_detect": [C.POINTER(DetectDesc), vp],
        "sbod_loss_forward_host": [C.POINTER(LossDesc), i32, vp, vp, sz, vp],
        "sbod_fcos_forward": [C.POINTER(FcosDesc), vp],
        "sbod_fcos_backward": [C.POINTER(FcosDesc), vp, vp, vp, vp, vp],
        "sbod_fcos_postprocess": [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name, args in {
        "sbod_nms_workspace_bytes": [i32],
        "sbod_assign_workspace_bytes": [i32, i32],
        "sbod_match_workspace_bytes": [i32, i32],
        "sbod_loss_workspace_bytes": [C.POINTER(LossDesc)],
        "sbod_detect_workspace_bytes": [C.POINTER(DetectDesc)],
        "sbod_detect_workspace_zero_bytes": [C.POINTER(DetectDesc)],
        "sbod_loss_forward_host_arena_bytes": [C.POINTER(LossDesc), i32],
        "sbod_fcos_workspace_bytes": [C.POINTER(FcosDesc)],
        "sbod_map_workspace_bytes": [i32],
    }.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_size_t


def lib():
    """Load libsbod.so (once). Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise SbodError(
                        f"{LIB_PATH} not found: build it with __graft_entry__.build() or "
                        "shape_based_object_detection_b200/csrc/build.sh (no CPU fallback exists)")
                handle = C.CDLL(LIB_PATH)
                _declare(handle)
                if handle.sbod_abi_version() != 1:
                    raise SbodError("libsbod.so ABI version mismatch")
                _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise SbodError(f"libsbod: {lib().sbod_error_string(rc).decode()} (code {rc})")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SbodError("sbod operators run on CUDA tensors only (no CPU fallback); got a "
                            f"{t.device} tensor")


def f32c(t):
    """fp32, contiguous, 16-byte aligned view/copy of t."""
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


class Workspace:
    """Grow-only, zero-initialised device scratch, one per (device, tag)."""

    _cache = {}

    @classmethod
    def get(cls, device, tag, nbytes, zero_bytes=None):
        key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
            # torch's allocator returns >=512-byte aligned blocks
            z = nbytes if zero_bytes is None else zero_bytes
            check(lib().sbod_workspace_init(ptr(buf), C.c_size_t(int(z)), stream_ptr()))
            cls._cache[key] = buf
        return buf
