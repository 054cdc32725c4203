"""Shared body of the anchor-based loss modules (MultiBoxLoss300/512, RetinaFocalLoss, RefineDetLoss).

The reference repeats ~120 lines per class (models/SSD300.py:446-594, SSD512.py:477-626,
RetinaNet.py:353-506, RefineDet512.py:698-956); here each class is a small spec that selects the
variant of the fused CUDA path (SURVEY.md §8 a' variant matrix)."""
import torch
import torch.nn as nn

from .. import _lib as L
from ..core import LossSpec, fused_loss


def _cfg(config, key, default=None):
    try:
        return getattr(config, key)
    except (AttributeError, KeyError):  # attribute-style dicts raise KeyError from __getattr__
        try:
            return config[key]
        except (KeyError, TypeError, IndexError):
            return default


class FusedAnchorLoss(nn.Module):
    """Base: holds priors (cxcy + the xyxy form the reference precomputes in its ctor) and config."""

    #: loc loss used when config.reg_loss != 'DIOU'
    plain_reg_kind = L.REG_SMOOTH_L1
    #: (mining variant for the CE branch, focal variant for the FOCAL branch)
    ce_kind = L.CLS_CE_MINE_NONPOS
    focal_kind = L.CLS_FOCAL_SUM
    #: The reference only special-cases reg_loss == 'DIOU' (SSD512.py:578-583, RetinaNet.py:461-466): any other
    #: value, 'GIOU' included, trains with the plain loc loss. Setting this attribute (or the config key
    #: 'sbod_extended_reg_losses') to True is the explicit opt-in that lets 'GIOU' / 'IOU' / 'CIOU' select the
    #: corresponding IouLoss flavour of operators/Loss.py:164-200 on the fused path (BASELINE config 3).
    extended_reg_losses = False

    def __init__(self, priors_cxcy, config, threshold=0.5, neg_pos_ratio=3):
        super().__init__()
        L.need_cuda(priors_cxcy)
        self.priors_cxcy = L.f32c(priors_cxcy.detach())
        from ..dataset.transforms import cxcy_to_xy
        self.priors_xy = cxcy_to_xy(self.priors_cxcy)  # same fp32 arithmetic as SSD512.py:488
        self.threshold = threshold
        self.neg_pos_ratio = neg_pos_ratio
        self.alpha = _cfg(config, "reg_weights", 1.0)
        self.device = _cfg(config, "device")
        self.n_classes = _cfg(config, "n_classes")
        self.config = config
        self.process_group = None  # set to a torch.distributed group to shard the batch by image
        self.last = {}             # device state of the last forward (for tests / inspection)

    def increase_threshold(self, increment=0.1):
        if self.threshold >= 0.7:
            return
        self.threshold += increment

    def _spec(self):
        reg = str(_cfg(self.config, "reg_loss", "")).upper()
        cls = str(_cfg(self.config, "cls_loss", "")).upper()
        kinds = {"DIOU": L.REG_DIOU}
        if self.extended_reg_losses or _cfg(self.config, "sbod_extended_reg_losses", False):
            kinds.update({"GIOU": L.REG_GIOU, "IOU": L.REG_IOU, "CIOU": L.REG_CIOU})
        reg_kind = kinds.get(reg, self.plain_reg_kind)
        cls_kind = self.focal_kind if cls == "FOCAL" else self.ce_kind
        return LossSpec(reg_kind=reg_kind, cls_kind=cls_kind, threshold=self.threshold,
                        neg_pos_ratio=self.neg_pos_ratio, reg_weight=float(self.alpha))

    def forward(self, predicted_locs, predicted_scores, boxes, labels):
        self.last = {}
        return fused_loss(self._spec(), self.priors_cxcy, self.priors_xy, predicted_locs, predicted_scores,
                          boxes, labels, group=self.process_group, lane=getattr(self, "exchange_lane", 0), holder=self.last)

    def forward_packed(self, predicted_locs, predicted_scores, packed_gt):
        """Same loss with the ground truth already packed on the device (core.pack_ground_truth):
        no host-side work, so the call can be captured in a CUDA graph."""
        self.last = {}
        return fused_loss(self._spec(), self.priors_cxcy, self.priors_xy, predicted_locs, predicted_scores,
                          None, None, group=self.process_group, lane=getattr(self, "exchange_lane", 0), holder=self.last, packed_gt=packed_gt)
