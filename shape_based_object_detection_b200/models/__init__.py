from .FCOSDet import FCOSLoss, compute_location, postprocess as fcos_postprocess
from .RefineDet512 import RefineDetLoss, offset2bbox
from .RetinaNet import RetinaFocalLoss
from .SSD300 import MultiBoxLoss300
from .SSD512 import MultiBoxLoss512
