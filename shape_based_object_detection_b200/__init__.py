"""B200-native (sm_100a) detection box pipeline — drop-in for the hot path of
shuaiqi361/shape_based_object_detection.

The sub-modules mirror the reference's module paths for that path only:
  metrics.find_jaccard_overlap / calculate_mAP, eval_results.coco_format_results, operators.iou_utils, operators.Loss, dataset.transforms (box
  converters), models.{SSD300,SSD512,RetinaNet,RefineDet512} (loss classes), models.utils.detect,
  detect_scripts.detect_tools.
All numerics run in hand-written CUDA kernels behind the C ABI of include/sbod.h (lib/libsbod.so).
"""
from . import _lib, core, eval_results
from .core import (LossSpec, assign, detect_batched, detect_begin, detect_end, fused_loss, pack_ground_truth,
                   unpad_detections)

__all__ = ["LossSpec", "assign", "detect_batched", "detect_begin", "detect_end", "fused_loss", "pack_ground_truth", "unpad_detections",
           "install", "_lib"]


def install(reference_root=None):
    """Patch the reference's modules in place so its drivers (train_anchor.py, train_refine.py,
    eval.py, detect_bboxes.py) pick up the CUDA path without edits. See INTEGRATION.md."""
    from .compat import install as _install
    return _install(reference_root)
