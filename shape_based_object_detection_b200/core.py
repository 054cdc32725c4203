"""Host-side drivers of the fused kernels: GT packing, workspaces, autograd glue.

Everything numerical happens in libsbod.so; this module only moves pointers.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L


# ---------------------------------------------------------------------------------------------
# GT packing: the reference passes python lists of per-image tensors (dataset/Datasets.py:58-86)
# ---------------------------------------------------------------------------------------------
def pack_ground_truth(boxes, labels, device):
    """list[N] of [G_i,4] / [G_i]  ->  (boxes [T,4] f32, labels [T] i64, offsets [N+1] i32, gmax).
    Host lists are packed into one pinned buffer and moved with a single copy (dataset.collate.PackedGT)."""
    if len(boxes) and all((not b.is_cuda) for b in boxes) and torch.device(device).type == "cuda":
        from .dataset.collate import PackedGT
        return PackedGT.from_lists(boxes, labels).to(device).as_tuple()
    counts = [int(b.shape[0]) for b in boxes]
    offs = np.zeros(len(counts) + 1, dtype=np.int32)
    np.cumsum(counts, out=offs[1:])
    if sum(counts) == 0:
        gt_b = torch.zeros((1, 4), dtype=torch.float32, device=device)
        gt_l = torch.zeros((1,), dtype=torch.int64, device=device)
    else:
        gt_b = torch.cat([b.reshape(-1, 4) for b in boxes], 0).to(device=device, dtype=torch.float32).contiguous()
        gt_l = torch.cat([l.reshape(-1) for l in labels], 0).to(device=device, dtype=torch.int64).contiguous()
    gt_o = torch.from_numpy(offs).to(device, non_blocking=True)
    return gt_b, gt_l, gt_o, (max(counts) if counts else 0)


@dataclass
class LossSpec:
    reg_kind: int
    cls_kind: int
    threshold: float = 0.5
    neg_pos_ratio: int = 3
    reg_weight: float = 1.0
    beta: float = 1.0 / 9.0
    focal_alpha: float = 0.25
    focal_gamma: float = 2.0
    binarize: bool = False
    neg_margin: float = 0.1


class LossState:
    """Per-call device state of one fused loss evaluation (kept alive for backward / inspection)."""

    def __init__(self, spec, priors_cxcy, priors_xy, locs, scores, gt, anchors_xy=None, exclude=None,
                 prefill_grad=False, comm=None):
        self.spec = spec
        dev = scores.device
        N, P, Cn = scores.shape
        self.N, self.P, self.C = N, P, Cn
        self.locs, self.scores = locs, scores
        self.priors_cxcy, self.priors_xy = priors_cxcy, priors_xy
        self.gt_boxes, self.gt_labels, self.gt_offsets, self.gmax = gt
        self.anchors_xy, self.exclude = anchors_xy, exclude
        self.ov = torch.empty((N, P), dtype=torch.float32, device=dev)
        self.obj = torch.empty((N, P), dtype=torch.int32, device=dev)
        self.lse = torch.empty((N, P), dtype=torch.float32, device=dev)
        self.ce = torch.empty((N, P), dtype=torch.float32, device=dev)
        self.sel = torch.empty((N, P), dtype=torch.uint8, device=dev)
        self.sel_thr = torch.empty((N, 2), dtype=torch.float32, device=dev)
        self.partials = torch.empty((N, 4), dtype=torch.float64, device=dev)
        self.sums = torch.empty((4,), dtype=torch.float64, device=dev)
        self.loss = torch.empty((4,), dtype=torch.float32, device=dev)
        d = L.LossDesc()
        d.locs, d.scores = locs.data_ptr(), scores.data_ptr()
        d.priors_cxcy, d.priors_xy = priors_cxcy.data_ptr(), priors_xy.data_ptr()
        d.anchors_xy = anchors_xy.data_ptr() if anchors_xy is not None else None
        d.gt_boxes, d.gt_labels, d.gt_offsets = (self.gt_boxes.data_ptr(), self.gt_labels.data_ptr(),
                                                 self.gt_offsets.data_ptr())
        d.exclude = exclude.data_ptr() if exclude is not None else None
        # objects per image change from batch to batch: the per-object tables are sized for the next power
        # of two, so a run sees a handful of workspace layouts instead of one per distinct maximum
        d.N, d.P, d.C, d.gmax = N, P, Cn, L.bucket(self.gmax)
        # python-float thresholds are compared in fp32 by torch (SURVEY §8a''): cast here
        d.thr_pos = float(np.float32(spec.threshold))
        d.thr_neg = float(np.float32(spec.threshold - spec.neg_margin))
        d.reg_kind, d.cls_kind = spec.reg_kind, spec.cls_kind
        d.binarize_labels = 1 if spec.binarize else 0
        d.neg_pos_ratio = int(spec.neg_pos_ratio)
        d.reg_weight, d.smooth_l1_beta = float(spec.reg_weight), float(spec.beta)
        d.focal_alpha, d.focal_gamma = float(spec.focal_alpha), float(spec.focal_gamma)
        d.ov, d.obj, d.lse, d.ce, d.sel = (self.ov.data_ptr(), self.obj.data_ptr(), self.lse.data_ptr(),
                                           self.ce.data_ptr(), self.sel.data_ptr())
        d.sel_thr = self.sel_thr.data_ptr()
        d.partials, d.sums, d.loss = self.partials.data_ptr(), self.sums.data_ptr(), self.loss.data_ptr()
        nbytes = L.lib().sbod_loss_workspace_bytes(C.byref(d))
        zbytes = L.lib().sbod_loss_workspace_zero_bytes(C.byref(d))
        self.ws = L.Workspace.get(dev, "loss", nbytes, zero_bytes=zbytes, layout=(N, P, d.gmax))
        d.workspace, d.workspace_bytes = self.ws.data_ptr(), nbytes
        # the gradient wrt the logits is zero almost everywhere: let the forward stream zero-fill it
        self.grad_scores = torch.empty_like(scores) if prefill_grad else None
        d.grad_scores_prefill = self.grad_scores.data_ptr() if prefill_grad else None
        # sharded batch: the forward's last kernel all-reduces `sums` itself over the NVLink mailboxes
        self.comm = comm
        d.comm = comm.ptr if comm is not None else None
        self.desc = d

    def forward(self):
        with torch.cuda.device(self.scores.device):
            L.check(L.lib().sbod_loss_forward(C.byref(self.desc), L.stream_ptr()))

    def finalize(self):
        with torch.cuda.device(self.scores.device):
            L.check(L.lib().sbod_loss_finalize(C.byref(self.desc), L.stream_ptr()))

    def backward(self, grad_loss, want_locs=True, want_scores=True):
        dev = self.scores.device
        g_locs = torch.empty_like(self.locs) if want_locs else None
        g_scores = None
        if want_scores:
            # the prefilled buffer is handed over (autograd can then keep it as .grad without a copy);
            # a second backward through the same state gets a fresh buffer and the full zero-fill
            g_scores, self.grad_scores = self.grad_scores, None
            if g_scores is None:
                g_scores = torch.empty_like(self.scores)
        gl = grad_loss.to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            L.check(L.lib().sbod_loss_backward(C.byref(self.desc), L.ptr(gl), L.ptr(g_locs), L.ptr(g_scores),
                                               L.stream_ptr()))
        return g_locs, g_scores

    def backward_into(self, grad_loss, g_locs, g_scores):
        """Backward into caller-owned buffers (no allocation)."""
        with torch.cuda.device(self.scores.device):
            L.check(L.lib().sbod_loss_backward(C.byref(self.desc), L.ptr(grad_loss), L.ptr(g_locs), L.ptr(g_scores),
                                               L.stream_ptr()))
        return g_locs, g_scores

    def targets(self):
        """(true_classes, true_neg_classes) as the reference materialises them — for tests."""
        dev = self.scores.device
        cls = torch.empty((self.N, self.P), dtype=torch.int64, device=dev)
        neg = torch.empty((self.N, self.P), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().sbod_loss_targets(C.byref(self.desc), L.ptr(cls), L.ptr(neg), L.stream_ptr()))
        return cls, neg


def _all_reduce_sums(state, group):
    import torch.distributed as dist
    dist.all_reduce(state.sums, op=dist.ReduceOp.SUM, group=group)
    state.finalize()


class _FusedLossFn(torch.autograd.Function):
    """loss = conf + alpha * loc, forward and backward entirely in libsbod kernels."""

    @staticmethod
    def forward(ctx, locs, scores, holder):
        state = holder["make_state"](locs, scores)
        state.forward()
        if holder.get("group") is not None and state.comm is None:
            _all_reduce_sums(state, holder["group"])  # NCCL path (peer exchange disabled or unavailable)
        holder["state"] = state
        ctx.state = state
        return state.loss[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        g_locs, g_scores = ctx.state.backward(grad_out, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return g_locs, g_scores, None


@L.on_device
def fused_loss(spec, priors_cxcy, priors_xy, predicted_locs, predicted_scores, boxes, labels,
               anchors_xy=None, exclude=None, group=None, holder=None, packed_gt=None, lane=0):
    """Run the fused train path. Returns a 0-dim fp32 tensor with grad_fn.
    packed_gt: optional result of pack_ground_truth() (CSR ground truth already on the device, e.g.
    packed by the data loader) — then `boxes` / `labels` are ignored and no host work is done."""
    L.need_cuda(predicted_locs, predicted_scores, priors_cxcy)
    dev = predicted_scores.device
    if predicted_locs.dim() != 3 or predicted_scores.dim() != 3:
        raise ValueError("predicted_locs / predicted_scores must be [N,P,4] / [N,P,C]")
    n_priors = priors_cxcy.size(0)
    assert n_priors == predicted_locs.size(1) == predicted_scores.size(1)  # SSD512.py:523
    from .dataset.collate import PackedGT
    if packed_gt is None and isinstance(boxes, PackedGT):  # the whole batch in one buffer (dataset.collate)
        packed_gt = (boxes if boxes.device == dev else boxes.to(dev)).as_tuple()
    gt = packed_gt if packed_gt is not None else pack_ground_truth(boxes, labels, dev)
    if group is not None and spec.cls_kind == L.CLS_CE_MINE_BATCH:
        # SSD300 mines its hard negatives over the WHOLE batch (SSD300.py:580-588): a rank's own top-k cannot
        # be rebuilt from per-rank sums, so this variant does not shard ("replicas only")
        raise L.SbodError("MultiBoxLoss300's batch-global hard-negative mining cannot be sharded by image: "
                          "run it with process_group=None (one replica per GPU)")
    holder = holder if holder is not None else {}
    holder["group"] = group
    # decided here: inside autograd.Function.forward grad mode is off. With CE + mining the gradient of
    # the logits is zero except on a few rows, and the forward's streaming kernel zero-fills it on the way.
    prefill = bool(predicted_scores.requires_grad and torch.is_grad_enabled()
                   and spec.cls_kind in (L.CLS_CE_MINE_NONPOS, L.CLS_CE_MINE_NEG, L.CLS_CE_MINE_BATCH))

    comm = None
    if group is not None:
        from .parallel import peer_exchange
        comm = peer_exchange(group, dev, lane)

    def make_state(locs, scores):
        return LossState(spec, priors_cxcy, priors_xy, L.f32c(locs.detach()), L.f32c(scores.detach()), gt,
                         anchors_xy=anchors_xy, exclude=exclude, prefill_grad=prefill, comm=comm)

    holder["make_state"] = make_state
    return _FusedLossFn.apply(predicted_locs, predicted_scores, holder)


# ---------------------------------------------------------------------------------------------
# stand-alone batched assignment
# ---------------------------------------------------------------------------------------------
@L.on_device
def assign(boxes, labels, anchors_xy, threshold=0.5, neg_margin=0.1, want_classes=True):
    """Batched assignment; anchors_xy [P,4] or [N,P,4]. Returns ov, obj(int32), true_classes, true_neg."""
    L.need_cuda(anchors_xy)
    dev = anchors_xy.device
    gt_b, gt_l, gt_o, gmax = pack_ground_truth(boxes, labels, dev)
    N = len(boxes)
    per_image = anchors_xy.dim() == 3
    P = anchors_xy.size(-2)
    a = L.f32c(anchors_xy)
    ov = torch.empty((N, P), dtype=torch.float32, device=dev)
    obj = torch.empty((N, P), dtype=torch.int32, device=dev)
    cls = torch.empty((N, P), dtype=torch.int64, device=dev) if want_classes else None
    neg = torch.empty((N, P), dtype=torch.int64, device=dev) if want_classes else None
    gmax = L.bucket(gmax)
    nbytes = L.lib().sbod_assign_workspace_bytes(N, gmax)
    ws = L.Workspace.get(dev, "assign", nbytes, layout=(N, gmax))
    L.check(L.lib().sbod_assign(L.ptr(gt_b), L.ptr(gt_l), L.ptr(gt_o), N, gmax, L.ptr(a),
                                1 if per_image else 0, P, float(np.float32(threshold)),
                                float(np.float32(threshold - neg_margin)), L.ptr(ov), L.ptr(obj),
                                L.ptr(cls), L.ptr(neg), L.ptr(ws), C.c_size_t(nbytes), L.stream_ptr()))
    return ov, obj, cls, neg


# ---------------------------------------------------------------------------------------------
# eval path
# ---------------------------------------------------------------------------------------------
@L.on_device
def make_detect_desc(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy=None,
                     act="softmax", box_type="offset", clamp_inplace=False, prior_keep=None,
                     second_nms_thr=-1.0, pre_nms_topk=0, class_agnostic=False):
    """Allocate the outputs / workspace of one sbod_detect call and fill its descriptor."""
    L.need_cuda(predicted_locs, predicted_scores)
    dev = predicted_scores.device
    N, P, Cn = predicted_scores.shape
    if clamp_inplace:
        if not (predicted_locs.is_contiguous() and predicted_locs.dtype == torch.float32
                and predicted_locs.data_ptr() % 16 == 0):
            raise L.SbodError("in-place clamp needs a contiguous, 16-byte aligned fp32 predicted_locs")
        locs = predicted_locs
    else:
        locs = L.f32c(predicted_locs.detach())
    scores = L.f32c(predicted_scores.detach())
    box_kind = {"offset": L.BOX_OFFSET, "center": L.BOX_CENTER}.get(box_type, L.BOX_CORNER)
    pri = L.f32c(priors_cxcy) if (priors_cxcy is not None and box_kind == L.BOX_OFFSET) else None
    keep = None
    if prior_keep is not None:
        keep = prior_keep.to(device=dev, dtype=torch.uint8).contiguous()
    cap = max(int(top_k), 1)
    out_boxes = torch.empty((N, cap, 4), dtype=torch.float32, device=dev)
    out_labels = torch.empty((N, cap), dtype=torch.int64, device=dev)
    out_scores = torch.empty((N, cap), dtype=torch.float32, device=dev)
    out_prior = torch.empty((N, cap), dtype=torch.int32, device=dev)
    out_counts = torch.empty((N,), dtype=torch.int32, device=dev)
    d = L.DetectDesc()
    d.locs, d.scores = locs.data_ptr(), scores.data_ptr()
    d.priors_cxcy = pri.data_ptr() if pri is not None else None
    d.prior_keep = keep.data_ptr() if keep is not None else None
    d.N, d.P, d.C = N, P, Cn
    d.act_kind = {"sigmoid": L.ACT_SIGMOID, "none": L.ACT_NONE}.get(act, L.ACT_SOFTMAX)
    d.box_kind, d.clamp_inplace = box_kind, 1 if clamp_inplace else 0
    d.min_score, d.max_overlap, d.top_k = float(min_score), float(max_overlap), int(top_k)
    d.second_nms_thr, d.pre_nms_topk = float(second_nms_thr), int(pre_nms_topk)
    d.class_agnostic = 1 if class_agnostic else 0
    d.out_boxes, d.out_labels, d.out_scores = out_boxes.data_ptr(), out_labels.data_ptr(), out_scores.data_ptr()
    d.out_prior, d.out_counts, d.out_cap = out_prior.data_ptr(), out_counts.data_ptr(), cap
    nbytes = L.lib().sbod_detect_workspace_bytes(C.byref(d))
    zbytes = L.lib().sbod_detect_workspace_zero_bytes(C.byref(d))
    ws = L.Workspace.get(dev, "detect", nbytes, zero_bytes=zbytes, layout=(N, Cn))
    d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
    return {"desc": d, "outputs": (out_boxes, out_labels, out_scores, out_prior, out_counts),
            "alive": (locs, scores, pri, keep, ws)}


@L.on_device
def detect_batched(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy=None,
                   act="softmax", box_type="offset", clamp_inplace=False, prior_keep=None,
                   second_nms_thr=-1.0, pre_nms_topk=0, class_agnostic=False):
    """Fused eval path. Returns padded outputs (boxes [N,K,4], labels [N,K], scores [N,K],
    prior [N,K], counts [N]) — all on the device, no host synchronisation.
    class_agnostic: the detect_objects variant (one candidate per prior, one class-agnostic NMS)."""
    call = make_detect_desc(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy, act,
                            box_type, clamp_inplace, prior_keep, second_nms_thr, pre_nms_topk, class_agnostic)
    L.check(L.lib().sbod_detect(C.byref(call["desc"]), L.stream_ptr()))
    return call["outputs"]


@L.on_device
def detect_begin(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy=None,
                 act="softmax", box_type="offset", clamp_inplace=False, prior_keep=None,
                 second_nms_thr=-1.0, pre_nms_topk=0, side_stream=None):
    """Split-phase detect_batched: enqueue the bound pass (the kernel that streams the logits) on
    `side_stream` now, so that it overlaps whatever the caller runs on the current stream before
    detect_end(). Same arguments / results as detect_batched."""
    call = make_detect_desc(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy, act,
                            box_type, clamp_inplace, prior_keep, second_nms_thr, pre_nms_topk)
    cur = torch.cuda.current_stream(predicted_scores.device)
    side = side_stream if side_stream is not None else cur
    if side is not cur:
        side.wait_stream(cur)
    with torch.cuda.stream(side):
        L.check(L.lib().sbod_detect_stage(C.byref(call["desc"]), 2, L.stream_ptr()))
    call["side"] = side if side is not cur else None
    return call


def detect_end(call):
    """Second half of detect_begin: exact evaluation of the candidate rows + NMS on the current stream."""
    with torch.cuda.device(call["outputs"][0].device):
        if call.get("side") is not None:
            torch.cuda.current_stream().wait_stream(call["side"])
        L.check(L.lib().sbod_detect_stage(C.byref(call["desc"]), 4, L.stream_ptr()))
    return call["outputs"]


def unpad_detections(out_boxes, out_labels, out_scores, out_counts):
    """Padded device outputs -> the reference's three python lists of per-image tensors.
    One D2H read of the N counts (the reference syncs N*(C-1) times, models/utils.py:252)."""
    counts = out_counts.tolist()
    if any(c < 0 for c in counts):
        raise L.SbodError("sbod_detect: kept-list capacity exceeded (code %s): the first NMS stage of detect_tools "
                          "kept more than 65536 boxes in one image" % min(counts))
    boxes = [out_boxes[i, :c] for i, c in enumerate(counts)]
    labels = [out_labels[i, :c] for i, c in enumerate(counts)]
    scores = [out_scores[i, :c] for i, c in enumerate(counts)]
    return boxes, labels, scores
