"""FCOSLoss, compute_location and postprocess of the reference's models/FCOSDet.py on CUDA kernels.

The reference versions do not run (SURVEY.md §8 a-F: the constructor, get_sample_region,
assign_targets, forward and postprocess each raise). This module keeps their names, signatures and
constants (FCOSDet.py:235-270, 311-544) and implements the semantics that file specifies, with the
canonical FCOS meaning where a line cannot execute:
  * one label/box target per location: among the objects whose centre box (half-size
    stride*radius, clipped to the object) contains the location and whose largest l/t/r/b distance
    lies in the level's size-of-interest range, the one of minimum area (:370-421, :424-474);
  * centerness target sqrt(min(l,r)/max(l,r) * min(t,b)/max(t,b)) (:479-486);
  * loss = SigmoidFocalLoss/(n_pos + N) + reg_weights * centerness-weighted DIoU + BCEWithLogits
    (:527-544), the DIoU taken between the boxes the predicted / target distances span around the
    location (the reference hands the raw distances to a corner-box DIoU, :537).
PARITY UNPINNED with respect to the reference; pinned only to the corrected CPU restatement kept with the tests.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from ..core import pack_ground_truth

FMAP_DIMS = (64, 32, 16, 8, 4)                                         # FCOSDet.py:176 (512 input)
FPN_STRIDES = (8 / 512., 16 / 512., 32 / 512., 64 / 512., 128 / 512.)  # :333
SIZES = ((0., 0.08), (0.08, 0.16), (0.16, 0.32), (0.32, 0.64), (0.64, 1.))  # :334
RADIUS = 1.5                                                           # :335


PIXEL_STRIDES = (8, 16, 32, 64, 128)


def compute_location(fmap_dims=FMAP_DIMS, device="cpu", image_size=None):
    """List (one tensor [cells, 2] per pyramid level) of cell centres, FCOSDet.py:235-251 (square 512 input:
    d x d cells at ((j + .5) / d, (i + .5) / d)). image_size=(H, W) gives the rectangular generalisation the
    reference's hard-coded 512 grid lacks (BASELINE config 5, 800 x 1333 -> 22 300 locations): level k has
    ceil(H / s_k) x ceil(W / s_k) cells centred at ((j + .5) s_k / W, (i + .5) s_k / H)."""
    if torch.device(device).type == "cuda":
        return _compute_location_device(fmap_dims, torch.device(device), image_size)
    out = []
    if image_size is None:
        for d in fmap_dims:
            idx = (np.arange(d, dtype=np.float64) + 0.5) / d
            cy, cx = np.meshgrid(idx, idx, indexing="ij")
            out.append(torch.tensor(np.stack([cx.ravel(), cy.ravel()], 1).astype(np.float32)).to(device))
        return out
    H, W = image_size
    for s in PIXEL_STRIDES:
        ys = (np.arange(-(-H // s), dtype=np.float64) + 0.5) * s / H
        xs = (np.arange(-(-W // s), dtype=np.float64) + 0.5) * s / W
        cy, cx = np.meshgrid(ys, xs, indexing="ij")
        out.append(torch.tensor(np.stack([cx.ravel(), cy.ravel()], 1).astype(np.float32)).to(device))
    return out


def _compute_location_device(fmap_dims, device, image_size):
    """compute_location on the GPU (sbod_prior_grid, centres only): bit-identical to the host tables."""
    if image_size is None:
        dims = [(d, d, 1.0, float(d), 1.0, float(d)) for d in fmap_dims]
    else:
        H, W = image_size
        dims = [(-(-H // s), -(-W // s), float(s), float(W), float(s), float(H)) for s in PIXEL_STRIDES]
    n = len(dims)
    rows = (C.c_int32 * n)(*[d[0] for d in dims])
    cols = (C.c_int32 * n)(*[d[1] for d in dims])
    nsh = (C.c_int32 * n)(*([0] * n))
    scale = (C.c_double * (4 * n))(*[v for d in dims for v in d[2:]])
    total = sum(d[0] * d[1] for d in dims)
    out = torch.empty((total, 2), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        L.check(L.lib().sbod_prior_grid(n, rows, cols, nsh, scale, None, 0, L.ptr(out), C.c_longlong(total),
                                        L.stream_ptr()))
    return list(out.split([d[0] * d[1] for d in dims], 0))


def _cfg(config, key, default=None):
    try:
        return getattr(config, key)
    except (AttributeError, KeyError):  # attribute-style dicts raise KeyError from __getattr__
        try:
            return config[key]
        except (KeyError, TypeError, IndexError):
            return default


class _FcosFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, locs, scores, centerness, mod, gt):
        dev = scores.device
        N, P, Cn = scores.shape
        l_, s_, c_ = L.f32c(locs.detach()), L.f32c(scores.detach()), L.f32c(centerness.detach())
        lab = torch.empty((N, P), dtype=torch.int32, device=dev)
        tgt = torch.empty((N, P, 4), dtype=torch.float32, device=dev)
        sums = torch.empty((6,), dtype=torch.float64, device=dev)
        loss = torch.empty((4,), dtype=torch.float32, device=dev)
        d = L.FcosDesc()
        d.locs, d.scores, d.centerness = l_.data_ptr(), s_.data_ptr(), c_.data_ptr()
        d.locations, d.loc_aux = mod.all_locations.data_ptr(), mod.loc_aux.data_ptr()
        d.gt_boxes, d.gt_labels, d.gt_offsets = gt[0].data_ptr(), gt[1].data_ptr(), gt[2].data_ptr()
        d.N, d.P, d.C = N, P, Cn
        d.center_sample = 1 if mod.center_sample else 0
        d.reg_weight, d.focal_alpha, d.focal_gamma = float(mod.alpha), 0.25, 2.0
        d.lab, d.tgt, d.sums, d.loss = lab.data_ptr(), tgt.data_ptr(), sums.data_ptr(), loss.data_ptr()
        comm = None
        if mod.process_group is not None:
            from ..parallel import peer_exchange
            comm = peer_exchange(mod.process_group, dev, getattr(mod, "exchange_lane", 0))
        d.comm = comm.ptr if comm is not None else None
        nbytes = L.lib().sbod_fcos_workspace_bytes(C.byref(d))
        ws = L.Workspace.get(dev, "fcos", nbytes, zero_bytes=0)
        d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
        with torch.cuda.device(dev):
            L.check(L.lib().sbod_fcos_forward(C.byref(d), L.stream_ptr()))
            if mod.process_group is not None and comm is None:
                # sharded by image: [focal, sum((1-diou)*w), sum(w), bce, n_pos, n_images] is all that crosses GPUs
                import torch.distributed as dist
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=mod.process_group)
                L.check(L.lib().sbod_fcos_finalize(C.byref(d), L.stream_ptr()))
        ctx.keep = (d, l_, s_, c_, lab, tgt, sums, loss, ws, gt)
        mod.last = {"labels": lab, "targets": tgt, "sums": sums, "loss": loss}
        return loss[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        d, l_, s_, c_ = ctx.keep[:4]
        gl = grad_out.to(device=s_.device, dtype=torch.float32).reshape(1).contiguous()
        g_l = torch.empty_like(l_) if ctx.needs_input_grad[0] else None
        g_s = torch.empty_like(s_) if ctx.needs_input_grad[1] else None
        g_c = torch.empty_like(c_) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(s_.device):
            L.check(L.lib().sbod_fcos_backward(C.byref(d), L.ptr(gl), L.ptr(g_l), L.ptr(g_s), L.ptr(g_c),
                                               L.stream_ptr()))
        return g_l, g_s, g_c, None, None


class FCOSLoss(nn.Module):
    """FCOSLoss(locations, config, threshold=0.5, center_sample=True) — FCOSDet.py:311-544."""

    def __init__(self, locations, config, threshold=0.5, center_sample=True, image_size=None):
        """image_size=(H, W): locations come from compute_location(image_size=...) and the strides are
        s / W in x and s / H in y; default: the reference's square 512 grid (strides s / 512)."""
        super().__init__()
        self.threshold = threshold
        self.alpha = _cfg(config, "reg_weights", 1.0)
        self.device = _cfg(config, "device")
        self.n_classes = _cfg(config, "n_classes")
        self.config = config
        self.locations = locations
        self.center_sample = center_sample
        self.INF = 1e6
        self.fpn_strides = list(FPN_STRIDES)
        self.sizes = [list(s) for s in SIZES]
        self.radius = RADIUS
        L.need_cuda(*locations)
        self.all_locations = L.f32c(torch.cat(list(locations), 0))
        aux = []
        for lvl, loc in enumerate(locations):  # per location: sampling radius (x, y), size-of-interest range
            if image_size is None:
                sx = sy = np.float32(self.fpn_strides[lvl])
            else:
                sx = np.float32(PIXEL_STRIDES[lvl] / image_size[1])
                sy = np.float32(PIXEL_STRIDES[lvl] / image_size[0])
            row = torch.tensor([sx * np.float32(self.radius), sy * np.float32(self.radius),
                                self.sizes[lvl][0], self.sizes[lvl][1]], dtype=torch.float32)
            aux.append(row[None].expand(loc.size(0), 4))
        self.loc_aux = L.f32c(torch.cat(aux, 0).contiguous().to(self.all_locations.device))
        self.image_size = image_size
        self.process_group = None  # set to a torch.distributed group to shard the batch by image
        self.last = {}

    def increase_threshold(self, increment=0.1):  # FCOSDet.py:337-341
        if self.threshold >= 0.7:
            return
        self.threshold += increment

    def forward(self, predicted_locs, predicted_scores, predicted_centerness, boxes, labels):
        L.need_cuda(predicted_locs, predicted_scores, predicted_centerness)
        assert predicted_locs.size(1) == predicted_scores.size(1)  # FCOSDet.py:500
        from ..dataset.collate import PackedGT
        if isinstance(boxes, PackedGT):
            gt = (boxes if boxes.device == predicted_scores.device else boxes.to(predicted_scores.device)).as_tuple()
        else:
            gt = pack_ground_truth(boxes, labels, predicted_scores.device)
        return _FcosFn.apply(predicted_locs, predicted_scores, predicted_centerness, self, gt)


def postprocess(box_pred, cls_pred, center_pred, locations):
    """FCOS.postprocess (FCOSDet.py:253-270): (xyxy boxes [N,P,4], class probabilities * centerness
    [N,P,C]); feed them to models.utils.detect with box_type 'corner' — the probabilities must not be
    passed through another activation (core.detect_batched(act='none'))."""
    L.need_cuda(box_pred, cls_pred, center_pred)
    b, c, z = L.f32c(box_pred.detach()), L.f32c(cls_pred.detach()), L.f32c(center_pred.detach())
    loc = L.f32c(torch.cat(list(locations), 0)) if isinstance(locations, (list, tuple)) else L.f32c(locations)
    N, P, Cn = c.shape
    out_l = torch.empty((N, P, 4), dtype=torch.float32, device=c.device)
    out_s = torch.empty((N, P, Cn), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        L.check(L.lib().sbod_fcos_postprocess(L.ptr(b), L.ptr(c), L.ptr(z), L.ptr(loc), N, P, Cn, L.ptr(out_l),
                                              L.ptr(out_s), L.stream_ptr()))
    return out_l, out_s
