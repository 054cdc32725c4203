"""MultiBoxLoss300 (reference models/SSD300.py:446-594) on the fused CUDA path."""
from .. import _lib as L
from ._multibox import FusedAnchorLoss


class MultiBoxLoss300(FusedAnchorLoss):
    """L1 (element mean, nn.L1Loss) or DIoU loc loss; CE with BATCH-GLOBAL hard-negative mining over
    the true_neg == -1 rows (SSD300.py:580-588), or un-normalised softmax focal (:557-563).
    Batch-global mining does not shard by image: keep process_group = None for this class."""
    plain_reg_kind = L.REG_L1_ELEM_MEAN
    ce_kind = L.CLS_CE_MINE_BATCH
    focal_kind = L.CLS_FOCAL_SUM
