"""RetinaFocalLoss (reference models/RetinaNet.py:353-506) on the fused CUDA path."""
from .. import _lib as L
from ._multibox import FusedAnchorLoss


class RetinaFocalLoss(FusedAnchorLoss):
    """SmoothL1 or DIoU loc loss; CE mining restricted to true_neg == -1 rows (RetinaNet.py:493), or
    softmax focal divided by the number of positives (:471-472)."""
    plain_reg_kind = L.REG_SMOOTH_L1
    ce_kind = L.CLS_CE_MINE_NEG
    focal_kind = L.CLS_FOCAL_NORM
