"""operators/Loss.py of the reference (focal_loss, SigmoidFocalLoss, FocalLoss, IouLoss,
SmoothL1Loss), same constructors and call signatures, computed by CUDA kernels with hand-written
backward passes. Final scalar reductions are sums over a few thousand rows and use torch.sum on
the device (no arithmetic of the path happens in eager PyTorch)."""
import torch
import torch.nn as nn

from .. import _lib as L
from .iou_utils import (bbox_overlaps_ciou, bbox_overlaps_diou, bbox_overlaps_giou, bbox_overlaps_iou,
                        decode)


class _RowLoss(torch.autograd.Function):
    """row_out[M] = per-row loss, grad wrt logits produced in the same kernel."""

    @staticmethod
    def forward(ctx, logits, target, kind, a0, a1, gamma):
        x = L.f32c(logits.detach())
        t = target.to(device=x.device, dtype=torch.int64).contiguous()
        M, Cn = x.shape
        row = torch.empty((M,), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if logits.requires_grad else None
        with torch.cuda.device(x.device):
            if kind == "softmax":
                L.check(L.lib().sbod_softmax_focal(L.ptr(x), L.ptr(t), M, Cn, a0, a1, gamma, L.ptr(row),
                                                   L.ptr(grad), L.stream_ptr()))
            elif kind == "bce":
                L.check(L.lib().sbod_bce_focal(L.ptr(x), L.ptr(t), M, Cn, a0, gamma, L.ptr(row),
                                               L.ptr(grad), L.stream_ptr()))
            else:
                L.check(L.lib().sbod_sigmoid_focal(L.ptr(x), L.ptr(t), M, Cn, a0, gamma, L.ptr(row),
                                                   L.ptr(grad), L.stream_ptr()))
        ctx.grad = grad
        return row

    @staticmethod
    def backward(ctx, grad_row):
        g = ctx.grad * grad_row.unsqueeze(1) if ctx.grad is not None else None
        return g, None, None, None, None, None


def focal_loss(y_pred, y_true, alpha=0.25, gamma=2., device='cuda:0'):
    """Softmax focal loss, summed (Loss.py:9-38). Foreground rows: alpha*(1-p_t)^g*(-log p_t);
    background rows (target 0): (1-alpha)*p_0^g*(-log p_0) — the reference's background weight."""
    if isinstance(alpha, (list, tuple)):
        fore_alpha, back_alpha = alpha[0], alpha[1]
    else:
        fore_alpha, back_alpha = alpha, 1 - alpha
    L.need_cuda(y_pred)
    if y_pred.shape[0] == 0:
        return y_pred.sum() * 0.0
    return _RowLoss.apply(y_pred, y_true, "softmax", float(fore_alpha), float(back_alpha), float(gamma)).sum()


class SigmoidFocalLoss(nn.Module):
    """Loss.py:41-80."""

    def __init__(self, gamma, alpha, config):
        super().__init__()
        self.gamma = gamma
        self.alpha = alpha
        self.device = config.device

    def forward(self, out, target):
        L.need_cuda(out)
        if out.shape[0] == 0:
            return out.sum() * 0.0
        return _RowLoss.apply(out, target, "sigmoid", float(self.alpha), 0.0, float(self.gamma)).sum()


class FocalLoss(nn.Module):
    """Loss.py:83-103 (imported by the reference, never instantiated — SURVEY §8 a12): sigmoid focal
    loss with BCE-with-logits over a one-hot target of all columns, clamped probabilities, summed."""

    def __init__(self, alpha=0.25, gamma=2):
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma

    def forward(self, pred_logits, targets):
        L.need_cuda(pred_logits)
        if pred_logits.shape[0] == 0:
            return pred_logits.sum() * 0.0
        return _RowLoss.apply(pred_logits, targets, "bce", float(self.alpha), 0.0, float(self.gamma)).sum()


class IouLoss(nn.Module):
    """Loss.py:164-200."""

    def __init__(self, pred_mode='Corner', reduce='mean', variances=None, losstype='Diou'):
        super(IouLoss, self).__init__()
        self.reduce = reduce
        self.pred_mode = pred_mode
        self.variances = variances
        self.loss = losstype

    def forward(self, loc_p, loc_t, prior_data=None, weights=None):
        num = loc_p.shape[0]
        if self.pred_mode == 'Center':
            assert prior_data is not None
            decoded_boxes = decode(loc_p, prior_data, self.variances)
        else:
            decoded_boxes = loc_p
        if self.loss == 'Iou':
            loss = 1.0 - bbox_overlaps_iou(decoded_boxes, loc_t)
        elif self.loss == 'Giou':
            loss = 1.0 - bbox_overlaps_giou(decoded_boxes, loc_t)
        elif self.loss == 'Diou':
            loss = 1.0 - bbox_overlaps_diou(decoded_boxes, loc_t)
        else:
            loss = 1.0 - bbox_overlaps_ciou(decoded_boxes, loc_t)
        if weights is not None and weights.sum() > 1e-6:
            return (loss * weights).sum() / weights.sum()
        if self.reduce == 'mean':
            return loss.sum() / num
        return loss.sum()


class _SmoothL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, beta):
        p, t = L.f32c(pred.detach()), L.f32c(target.detach())
        out = torch.empty_like(p)
        grad = torch.empty_like(p) if pred.requires_grad else None
        L.device_of(p, t)
        with torch.cuda.device(p.device):
            L.check(L.lib().sbod_smooth_l1(L.ptr(p), L.ptr(t), p.numel(), float(beta), L.ptr(out), L.ptr(grad),
                                           L.stream_ptr()))
        ctx.grad = grad
        return out

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None, None


class SmoothL1Loss(nn.Module):
    """Loss.py:203-226: 'mean' divides by the number of ROWS."""

    def __init__(self, beta=1.0 / 9.0, reduction='mean'):
        super().__init__()
        self.beta = beta
        self.reduction = reduction

    def forward(self, pred, target, weights=None):
        L.need_cuda(pred, target)
        num = pred.size(0)
        l1_loss = _SmoothL1.apply(pred, target, self.beta)
        if weights is not None and weights.sum() > 1e-6:
            assert pred.size(0) == target.size(0) == weights.size(0)
            return (l1_loss * weights).sum() / weights.sum()
        if self.reduction == 'mean':
            return l1_loss.sum() / num
        return l1_loss.sum()
