"""metrics.find_jaccard_overlap (reference metrics.py:208-252) on the GPU."""
import torch

from . import _lib as L

EPS = 1e-5


def find_jaccard_overlap(gt_boxes, anchors):
    """[k,4] x [n,4] (xyxy) -> [k,n] IoU with the reference's EPS and zero-box masking
    (zero-size GT rows -> 0, zero-size anchors -> -1; metrics.py:235-250)."""
    L.need_cuda(gt_boxes, anchors)
    k, n = gt_boxes.size(0), anchors.size(0)
    out = torch.empty((k, n), dtype=torch.float32, device=gt_boxes.device)
    if k and n:
        a, b = L.f32c(gt_boxes), L.f32c(anchors)
        L.check(L.lib().sbod_iou_matrix(L.ptr(a), k, L.ptr(b), n, L.IOU_METRICS, L.ptr(out), L.stream_ptr()))
    return out
