"""metrics.find_jaccard_overlap (reference metrics.py:208-252) and metrics.calculate_mAP
(metrics.py:8-145) on the GPU."""
import ctypes as C

import torch

from . import _lib as L

EPS = 1e-5


@L.on_device
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


def _csr(tensors, dev, dtype, width=None):
    """list of per-image tensors -> (concatenated tensor on dev, offsets int32 [N+1], max rows)."""
    sizes = [int(t.size(0)) for t in tensors]
    off = [0]
    for n in sizes:
        off.append(off[-1] + n)
    shape = (0,) if width is None else (0, width)
    flat = torch.cat([t.reshape((-1,) if width is None else (-1, width)) for t in tensors], 0) if tensors \
        else torch.zeros(shape)
    return (flat.to(device=dev, dtype=dtype).contiguous(), torch.tensor(off, dtype=torch.int32, device=dev),
            max(sizes) if sizes else 0)


def calculate_mAP(det_boxes, det_labels, det_scores, true_boxes, true_labels, true_difficulties, threshold,
                  label_map, device="cuda:0"):
    """Same arguments and results as the reference's metrics.calculate_mAP (metrics.py:8-145): lists
    with one tensor per image; returns ({class name: AP}, mAP), VOC07 11-point interpolation, greedy
    matching in descending-score order, "difficult" objects ignored. Detections with equal scores are
    taken in the order of the concatenated lists (the reference's sort leaves ties unspecified).
    One sbod_map call: per-image matching kernel, a device-wide radix sort, per-class AP kernel."""
    assert len(det_boxes) == len(det_labels) == len(det_scores) == len(true_boxes) == len(true_labels) == \
        len(true_difficulties)  # metrics.py:25-27
    dev = torch.device(device)
    if dev.type != "cuda":
        raise L.SbodError("calculate_mAP runs on a CUDA device (there is no CPU fallback)")
    n_classes = len(label_map)
    n_images = len(det_boxes)
    db, d_off, _ = _csr(det_boxes, dev, torch.float32, 4)
    dl, _, _ = _csr(det_labels, dev, torch.int64)
    ds, _, _ = _csr(det_scores, dev, torch.float32)
    tb, g_off, gmax = _csr(true_boxes, dev, torch.float32, 4)
    tl, _, _ = _csr(true_labels, dev, torch.int64)
    td, _, _ = _csr(true_difficulties, dev, torch.uint8)
    assert db.size(0) == dl.size(0) == ds.size(0) and tb.size(0) == tl.size(0) == td.size(0)  # metrics.py:38,50
    D, T = int(db.size(0)), int(tb.size(0))
    ap = torch.zeros((n_classes - 1,), dtype=torch.float32, device=dev)
    recall_thresholds = torch.arange(start=0, end=1.1, step=.1)  # metrics.py:128, fp32
    thr = (C.c_float * 11)(*[float(v) for v in recall_thresholds.tolist()])
    nbytes = L.lib().sbod_map_workspace_bytes(D)
    with torch.cuda.device(dev):
        ws = L.Workspace.get(dev, "map", nbytes, zero_bytes=0)
    with torch.cuda.device(dev):
        L.check(L.lib().sbod_map(L.ptr(db), L.ptr(dl), L.ptr(ds), L.ptr(d_off), D, L.ptr(tb), L.ptr(tl), L.ptr(td),
                                 L.ptr(g_off), T, n_images, gmax, n_classes, float(threshold), thr, L.ptr(ap),
                                 L.ptr(ws), C.c_size_t(nbytes), L.stream_ptr()))
    average_precisions = ap.cpu()
    mean_average_precision = average_precisions.mean().item()
    rev_label_map = {v: k for k, v in label_map.items()}
    return {rev_label_map[c + 1]: v for c, v in enumerate(average_precisions.tolist())}, mean_average_precision
