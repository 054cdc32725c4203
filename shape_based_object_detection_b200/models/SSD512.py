"""MultiBoxLoss512 (reference models/SSD512.py:477-626) on the fused CUDA path."""
from .. import _lib as L
from ._multibox import FusedAnchorLoss


class MultiBoxLoss512(FusedAnchorLoss):
    """SmoothL1(beta=1/9, row mean) or DIoU loc loss; CE with per-image hard-negative mining where
    only positives are excluded (SSD512.py:610-619), or un-normalised softmax focal (:587-593)."""
    plain_reg_kind = L.REG_SMOOTH_L1
    ce_kind = L.CLS_CE_MINE_NONPOS
    focal_kind = L.CLS_FOCAL_SUM
