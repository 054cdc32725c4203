"""RefineDetLoss + offset2bbox (reference models/RefineDet512.py:643-653, 698-956) on the CUDA path."""
import torch
import torch.nn as nn

from .. import _lib as L
from ..core import LossSpec, fused_loss
from ._multibox import _cfg


@L.on_device
def offset2bbox(arm_locs, odm_locs, priors_cxcy):
    """Two-stage decode ARM -> ODM -> xyxy for the whole batch (RefineDet512.py:643-653)."""
    L.need_cuda(arm_locs, odm_locs, priors_cxcy)
    a, o, p = L.f32c(arm_locs.detach()), L.f32c(odm_locs.detach()), L.f32c(priors_cxcy)
    out = torch.empty_like(a)
    L.check(L.lib().sbod_offset2bbox(L.ptr(a), L.ptr(o), L.ptr(p), L.ptr(out), a.size(0), a.size(1),
                                     L.stream_ptr()))
    return out


def decode_arm(arm_locs, priors_cxcy):
    """cxcy_to_xy(gcxgcy_to_cxcy(arm_locs[i], priors)) for every image (RefineDet512.py:850)."""
    zeros = torch.zeros_like(arm_locs)
    # ODM offsets of zero leave the ARM box unchanged up to the exp(0)*w product
    return offset2bbox(arm_locs, zeros, priors_cxcy)


class RefineDetLoss(nn.Module):
    def __init__(self, priors_cxcy, config, threshold=0.5, neg_pos_ratio=3, theta=0.01):
        super().__init__()
        L.need_cuda(priors_cxcy)
        from ..dataset.transforms import cxcy_to_xy
        self.priors_cxcy = L.f32c(priors_cxcy.detach())
        self.priors_xy = cxcy_to_xy(self.priors_cxcy)
        self.threshold = threshold
        self.neg_pos_ratio = neg_pos_ratio
        self.alpha = _cfg(config, "reg_weights", 1.0)
        self.device = _cfg(config, "device")
        self.n_classes = _cfg(config, "n_classes")
        self.config = config
        self.theta = theta
        self.process_group = None
        self.last_arm, self.last_odm = {}, {}

    def increase_threshold(self, increment=0.05):  # RefineDet512.py:724-728
        if self.threshold + increment >= 0.7:
            self.threshold = 0.7
        else:
            self.threshold += increment

    def _spec(self, binarize):
        return LossSpec(reg_kind=L.REG_SMOOTH_L1, cls_kind=L.CLS_CE_MINE_NONPOS, threshold=self.threshold,
                        neg_pos_ratio=self.neg_pos_ratio, reg_weight=float(self.alpha), binarize=binarize)

    def compute_arm_loss(self, arm_locs, arm_scores, boxes, labels):
        """Binary ARM loss against the static priors (RefineDet512.py:730-820)."""
        self.last_arm = {}
        return fused_loss(self._spec(True), self.priors_cxcy, self.priors_xy, arm_locs, arm_scores, boxes,
                          labels, group=self.process_group, lane=getattr(self, "exchange_lane", 0), holder=self.last_arm)

    def compute_odm_loss(self, arm_locs, arm_scores, odm_locs, odm_scores, boxes, labels):
        """ODM loss against the refined anchors (RefineDet512.py:822-939)."""
        assert self.priors_cxcy.size(0) == odm_locs.size(1) == odm_scores.size(1)
        anchors_xy = decode_arm(arm_locs.detach(), self.priors_cxcy)
        # easy negatives: softmax(arm)[..., 1] < theta   (RefineDet512.py:894-899)
        a_sc = L.f32c(arm_scores.detach())
        assert a_sc.size(2) == 2
        exclude = torch.empty(a_sc.shape[:2], dtype=torch.uint8, device=a_sc.device)
        with torch.cuda.device(a_sc.device):
            L.check(L.lib().sbod_arm_easy_negative(L.ptr(a_sc), a_sc.size(0) * a_sc.size(1), float(self.theta),
                                                   L.ptr(exclude), L.stream_ptr()))
        self.last_odm = {}
        return fused_loss(self._spec(False), self.priors_cxcy, self.priors_xy, odm_locs, odm_scores, boxes,
                          labels, anchors_xy=anchors_xy, exclude=exclude, group=self.process_group, lane=getattr(self, "exchange_lane", 0),
                          holder=self.last_odm)

    def forward(self, arm_locs, arm_scores, odm_locs, odm_scores, boxes, labels):
        arm_loss = self.compute_arm_loss(arm_locs, arm_scores, boxes, labels)
        odm_loss = self.compute_odm_loss(arm_locs.data.detach(), arm_scores.data.detach(), odm_locs,
                                         odm_scores, boxes, labels)
        return arm_loss + odm_loss
