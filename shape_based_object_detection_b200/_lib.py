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
    "sbod_box_encode", "sbod_box_decode", "sbod_box_op_bwd", "sbod_prior_grid", "sbod_offset2bbox", "sbod_arm_easy_negative", "sbod_pair_iou_fwd",
    "sbod_pair_iou_bwd", "sbod_smooth_l1", "sbod_softmax_focal", "sbod_sigmoid_focal",
    "sbod_nms_workspace_bytes", "sbod_nms", "sbod_assign_workspace_bytes", "sbod_assign",
    "sbod_match_workspace_bytes", "sbod_match", "sbod_loss_workspace_bytes", "sbod_loss_workspace_zero_bytes",
    "sbod_workspace_init", "sbod_set_option",
    "sbod_loss_forward", "sbod_loss_forward_stage", "sbod_detect_stage", "sbod_loss_finalize", "sbod_loss_backward", "sbod_loss_targets",
    "sbod_detect_workspace_bytes", "sbod_detect_workspace_zero_bytes", "sbod_detect", "sbod_detect_probabilities",
    "sbod_comm_handle_bytes", "sbod_comm_create", "sbod_comm_connect", "sbod_comm_device_ptr", "sbod_comm_allreduce",
    "sbod_comm_destroy",
    "sbod_loss_forward_host_arena_bytes", "sbod_loss_forward_host",
    "sbod_fcos_workspace_bytes", "sbod_fcos_forward", "sbod_fcos_finalize", "sbod_fcos_backward", "sbod_fcos_postprocess",
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
        ("sel", C.c_void_p), ("sel_thr", C.c_void_p), ("partials", C.c_void_p), ("sums", C.c_void_p), ("loss", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("grad_scores_prefill", C.c_void_p), ("comm", C.c_void_p),
    ]


class DetectDesc(C.Structure):
    _fields_ = [
        ("locs", C.c_void_p), ("scores", C.c_void_p), ("priors_cxcy", C.c_void_p),
        ("prior_keep", C.c_void_p),
        ("N", C.c_int32), ("P", C.c_int32), ("C", C.c_int32),
        ("act_kind", C.c_int32), ("box_kind", C.c_int32), ("clamp_inplace", C.c_int32),
        ("min_score", C.c_float), ("max_overlap", C.c_float), ("top_k", C.c_int32),
        ("second_nms_thr", C.c_float), ("pre_nms_topk", C.c_int32), ("class_agnostic", C.c_int32),
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
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("comm", C.c_void_p),
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
        "sbod_box_op_bwd": [i32, vp, vp, vp, vp, i32, f32, f32, vp],
        "sbod_prior_grid": [i32, vp, vp, vp, vp, vp, i32, vp, C.c_longlong, vp],
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
        "sbod_set_option": [i32, i32],
        "sbod_loss_forward": [C.POINTER(LossDesc), vp],
        "sbod_loss_finalize": [C.POINTER(LossDesc), vp],
        "sbod_loss_forward_stage": [C.POINTER(LossDesc), i32, vp],
        "sbod_detect_stage": [C.POINTER(DetectDesc), i32, vp],
        "sbod_loss_backward": [C.POINTER(LossDesc), vp, vp, vp, vp],
        "sbod_loss_targets": [C.POINTER(LossDesc), vp, vp, vp],
        "sbod_detect": [C.POINTER(DetectDesc), vp],
        "sbod_detect_probabilities": [vp, i32, i32, i32, i32, vp, vp],
        "sbod_comm_create": [i32, i32, C.POINTER(vp), vp],
        "sbod_comm_connect": [vp, vp],
        "sbod_comm_allreduce": [vp, vp, i32, vp],
        "sbod_comm_destroy": [vp],
        "sbod_loss_forward_host": [C.POINTER(LossDesc), i32, vp, vp, sz, vp],
        "sbod_fcos_forward": [C.POINTER(FcosDesc), vp],
        "sbod_fcos_finalize": [C.POINTER(FcosDesc), vp],
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
        "sbod_loss_workspace_zero_bytes": [C.POINTER(LossDesc)],
        "sbod_detect_workspace_bytes": [C.POINTER(DetectDesc)],
        "sbod_detect_workspace_zero_bytes": [C.POINTER(DetectDesc)],
        "sbod_loss_forward_host_arena_bytes": [C.POINTER(LossDesc), i32],
        "sbod_fcos_workspace_bytes": [C.POINTER(FcosDesc)],
        "sbod_map_workspace_bytes": [i32],
        "sbod_comm_handle_bytes": [],
    }.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_size_t


def _declare_ptr_returns(lib):
    lib.sbod_comm_device_ptr.argtypes = [C.c_void_p]
    lib.sbod_comm_device_ptr.restype = C.c_void_p


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
                _declare_ptr_returns(handle)
                if handle.sbod_abi_version() != 2:
                    raise SbodError("libsbod.so ABI version mismatch")
                _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise SbodError(f"libsbod: {lib().sbod_error_string(rc).decode()} (code {rc})")


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (default: the current device). Entry points
    run under on_device(), so "current device" is the device of the tensors they were handed."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SbodError("sbod operators run on CUDA tensors only (no CPU fallback); got a "
                            f"{t.device} tensor")


def _cuda_devices(obj, found):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            found.add(obj.device.index)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _cuda_devices(o, found)


def device_of(*objs):
    """The one CUDA device the tensors in `objs` (tensors, nested lists / tuples) live on, or None when
    there is no CUDA tensor among them. Mixed devices are rejected: every sbod_* kernel takes raw
    pointers and launches on a single device."""
    found = set()
    _cuda_devices(objs, found)
    if len(found) > 1:
        raise SbodError("sbod operators need all tensors on ONE CUDA device; got devices %s" % sorted(found))
    return torch.device("cuda", found.pop()) if found else None


def on_device(fn):
    """Decorator of the public entry points: run `fn` with the tensors' device as the current CUDA device,
    so that stream_ptr(), the allocations and every kernel launch of the call target that device
    (the reference picks cuda:1 when config.device == 1, train_anchor.py:66-67)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = device_of(args, tuple(kwargs.values()))
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def f32c(t):
    """fp32, contiguous, 16-byte aligned view/copy of t."""
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def bucket(n):
    """Next power of two >= max(n, 1): data-dependent sizes (objects per image, boxes per NMS call) are
    rounded up before they size a workspace, so a training run sees a handful of layouts, not one per batch."""
    n = max(int(n), 1)
    return 1 << (n - 1).bit_length()


class Workspace:
    """Grow-only device scratch, ONE buffer per (device, stream, tag).

    * Per stream: two calls of the same operator on two streams (the train and the eval half of a step,
      ARM and ODM criteria on side streams) never share scratch, and the zero-initialisation is enqueued
      on the stream that uses the buffer.
    * Zero contract (sbod.h): the leading `zero_bytes` of a workspace must be zero before the first call
      with a given layout; every call leaves them zero again. `layout` names what the carve-up depends
      on; when it changes (or the buffer grows) the zero region is cleared again on the current stream.
    """

    _cache = {}

    @classmethod
    def get(cls, device, tag, nbytes, zero_bytes=None, layout=None):
        index = device.index if device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(device)
        key = (index, stream.cuda_stream, tag)
        ent = cls._cache.get(key)
        z = int(nbytes if zero_bytes is None else zero_bytes)
        layout = () if layout is None else layout
        if ent is None or ent[0].numel() < int(nbytes) + 256:
            # torch's allocator returns >=512-byte aligned blocks
            buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
            ent = [buf, None]  # None never equals a layout: the fresh buffer is cleared below
            cls._cache[key] = ent
        if ent[1] != layout:
            if z > 0:
                check(lib().sbod_workspace_init(ptr(ent[0]), C.c_size_t(z), C.c_void_p(stream.cuda_stream)))
            ent[1] = layout
        return ent[0]
