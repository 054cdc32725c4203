"""operators/iou_utils.py of the reference, served by CUDA kernels (same names and signatures).

Paired IoU family (iou_utils.py:6-164), point_form/center_size/intersect/jaccard (:167-233),
match_ious/match (:236-321), encode/decode (:324-368), log_sum_exp (:371-379), nms/diounms (:385-530).
"""
import ctypes as C

import torch

from .. import _boxops as B
from .. import _lib as L


class _PairOverlap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, b1, b2, kind):
        a, b = L.f32c(b1.detach()), L.f32c(b2.detach())
        out = torch.empty((a.size(0),), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            L.check(L.lib().sbod_pair_iou_fwd(L.ptr(a), L.ptr(b), a.size(0), kind, L.ptr(out), L.stream_ptr()))
        ctx.save_for_backward(a, b)
        ctx.kind = kind
        return out

    @staticmethod
    def backward(ctx, grad_out):
        a, b = ctx.saved_tensors
        go = L.f32c(grad_out)
        g1 = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(a.device):
            L.check(L.lib().sbod_pair_iou_bwd(L.ptr(a), L.ptr(b), L.ptr(go), a.size(0), ctx.kind, L.ptr(g1),
                                              L.ptr(g2), L.stream_ptr()))
        return g1, g2, None


@L.on_device
def _pair(bboxes1, bboxes2, kind):
    rows, cols = bboxes1.shape[0], bboxes2.shape[0]
    if rows * cols == 0:  # the reference returns its (rows, cols) scratch tensor here (iou_utils.py:9-11)
        return torch.zeros((rows, cols))
    L.need_cuda(bboxes1, bboxes2)
    if rows != cols:
        raise ValueError("paired IoU needs equally many boxes on both sides (got %d and %d)" % (rows, cols))
    return _PairOverlap.apply(bboxes1, bboxes2, kind)


def bbox_overlaps_iou(bboxes1, bboxes2):
    return _pair(bboxes1, bboxes2, L.PAIR_IOU)


def bbox_overlaps_giou(bboxes1, bboxes2):
    return _pair(bboxes1, bboxes2, L.PAIR_GIOU)


def bbox_overlaps_diou(bboxes1, bboxes2):
    return _pair(bboxes1, bboxes2, L.PAIR_DIOU)


def bbox_overlaps_ciou(bboxes1, bboxes2):
    return _pair(bboxes1, bboxes2, L.PAIR_CIOU)


def point_form(boxes):
    """(cx, cy, w, h) -> (xmin, ymin, xmax, ymax); iou_utils.py:167-177."""
    return B.convert(boxes, L.BOX_CXCY_TO_XY)


def center_size(boxes):
    """(xmin, ymin, xmax, ymax) -> (cx, cy, w, h). The reference's version (iou_utils.py:180-189)
    raises on its malformed torch.cat call; this is the intended conversion."""
    return B.convert(boxes, L.BOX_XY_TO_CXCY)


@L.on_device
def jaccard(box_a, box_b):
    """Dense IoU [A,B] without EPS or masking (0/0 -> NaN); iou_utils.py:215-233."""
    L.need_cuda(box_a, box_b)
    A, B = box_a.size(0), box_b.size(0)
    out = torch.empty((A, B), dtype=torch.float32, device=box_a.device)
    if A and B:
        a, b = L.f32c(box_a), L.f32c(box_b)
        L.check(L.lib().sbod_iou_matrix(L.ptr(a), A, L.ptr(b), B, L.IOU_JACCARD, L.ptr(out), L.stream_ptr()))
    return out


@L.on_device
def intersect(box_a, box_b):
    """Intersection area [A,B]; iou_utils.py:192-212."""
    L.need_cuda(box_a, box_b)
    A, B = box_a.size(0), box_b.size(0)
    out = torch.empty((A, B), dtype=torch.float32, device=box_a.device)
    if A and B:
        a, b = L.f32c(box_a), L.f32c(box_b)
        L.check(L.lib().sbod_iou_matrix(L.ptr(a), A, L.ptr(b), B, L.IOU_INTERSECT, L.ptr(out), L.stream_ptr()))
    return out


def encode(matched, priors, variances):
    """iou_utils.py:324-345."""
    return B.encode(matched, priors, L.CODEC_IOU_UTILS, variances[0], variances[1])


def decode(loc, priors, variances):
    """iou_utils.py:349-368 (returns xyxy); differentiable with respect to loc (IouLoss 'Center' mode)."""
    return B.decode(loc, priors, L.CODEC_IOU_UTILS, variances[0], variances[1])


@L.on_device
def _match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx, encode_loc):
    L.need_cuda(truths, priors, labels)
    dev = priors.device
    G, P = truths.size(0), priors.size(0)
    t, p = L.f32c(truths), L.f32c(priors)
    lab = labels.to(torch.int64).contiguous()
    loc = torch.empty((P, 4), dtype=torch.float32, device=dev)
    conf = torch.empty((P,), dtype=torch.int64, device=dev)
    nbytes = L.lib().sbod_match_workspace_bytes(G, P)
    kbytes = (G * 8 + 255) // 256 * 256  # the per-object keys lead the workspace and carry the zero contract
    ws = L.Workspace.get(dev, "match", nbytes, zero_bytes=kbytes, layout=(kbytes,))
    v0, v1 = (float(variances[0]), float(variances[1])) if variances is not None else (0.1, 0.2)
    L.check(L.lib().sbod_match(float(threshold), L.ptr(t), G, L.ptr(p), P, v0, v1, L.ptr(lab),
                               1 if encode_loc else 0, L.ptr(loc), L.ptr(conf), L.ptr(ws),
                               C.c_size_t(nbytes), L.stream_ptr()))
    loc_t[idx] = loc   # fills the caller's tensors in place and returns None, like the reference
    conf_t[idx] = conf.to(conf_t.dtype)


def match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx):
    """iou_utils.py:280-321."""
    _match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx, True)


def match_ious(threshold, truths, priors, variances, labels, loc_t, conf_t, idx):
    """iou_utils.py:236-277 (loc_t[idx] receives the matched xyxy boxes)."""
    _match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx, False)


def log_sum_exp(x):
    """iou_utils.py:371-379 (global max shift)."""
    x_max = x.data.max()
    return torch.log(torch.sum(torch.exp(x - x_max), 1, keepdim=True)) + x_max


@L.on_device
def _nms_device(boxes, scores, overlap, top_k, diou_beta=None):
    L.need_cuda(boxes, scores)
    n = scores.size(0)
    b, s = L.f32c(boxes), L.f32c(scores)
    keep = torch.zeros((n,), dtype=torch.int64, device=scores.device)
    count = torch.zeros((1,), dtype=torch.int32, device=scores.device)
    nbytes = L.lib().sbod_nms_workspace_bytes(n)
    ws = L.Workspace.get(scores.device, "nms", nbytes, zero_bytes=0)
    if diou_beta is None:
        L.check(L.lib().sbod_nms(L.ptr(b), L.ptr(s), n, float(overlap), int(top_k), L.ptr(keep), L.ptr(count),
                                 L.ptr(ws), C.c_size_t(nbytes), L.stream_ptr()))
    else:
        L.check(L.lib().sbod_diou_nms(L.ptr(b), L.ptr(s), n, float(overlap), int(top_k), float(diou_beta),
                                      L.ptr(keep), L.ptr(count), L.ptr(ws), C.c_size_t(nbytes), L.stream_ptr()))
    return keep, count


def nms(boxes, scores, overlap=0.5, top_k=200):
    """iou_utils.py:385-450: (keep [n] zero padded, count). Suppresses IoU > overlap among the top_k
    best-scored boxes."""
    if boxes.numel() == 0:
        return scores.new(scores.size(0)).zero_().long()  # bare tensor, as the reference does (:398-399)
    keep, count = _nms_device(boxes, scores, overlap, top_k)
    return keep, int(count.item())


def diounms(boxes, scores, overlap=0.5, top_k=200, beta1=1.0):
    """iou_utils.py:453-530 as written (never called by the reference): greedy NMS among the top_k
    best-scored boxes with the criterion IoU - (d / c) ** beta1 <= overlap, d using the candidate's
    y2 where its centre was meant (:507). Returns (keep [n] zero padded, count)."""
    if boxes.numel() == 0:
        return scores.new(scores.size(0)).zero_().long()  # bare tensor, as the reference does (:467-468)
    keep, count = _nms_device(boxes, scores, overlap, top_k, diou_beta=beta1)
    return keep, int(count.item())


def torchvision_nms(boxes, scores, iou_threshold):
    """Drop-in for torchvision.ops.nms (models/utils.py:265): kept indices, score-descending."""
    if scores.numel() == 0:
        return torch.zeros((0,), dtype=torch.int64, device=scores.device)
    keep, count = _nms_device(boxes, scores, iou_threshold, 0)
    return keep[: int(count.item())]
