"""models/utils.py:detect and detect_objects of the reference, on the fused CUDA eval path."""
import torch

from .. import _lib as L
from ..core import detect_batched, unpad_detections


def _cfg(config, key, default=None):
    try:
        return config[key]
    except (KeyError, TypeError, IndexError):
        return getattr(config, key, default)


def detect(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy, config,
           prior_positives_idx=None):
    """Reference models/utils.py:181-297. Returns three python lists (boxes, labels, scores) of N
    per-image tensors: classes ascending with NMS order inside a class, or the top_k best by score
    when more than top_k boxes survive; a [0,0,1,1]/0/0.0 placeholder when nothing does.
    box_type other than 'offset'/'center' clamps the caller's predicted_locs in place (:224)."""
    box_type = _cfg(config, "model")["box_type"]
    focal_type = _cfg(config, "focal_type")
    act = "sigmoid" if str(focal_type).lower() == "sigmoid" else "softmax"
    clamp_inplace = box_type not in ("offset", "center")
    out = detect_batched(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy,
                         act=act, box_type=box_type, clamp_inplace=clamp_inplace,
                         prior_keep=prior_positives_idx)
    return unpad_detections(out[0], out[1], out[2], out[4])


def detect_objects(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy, config):
    """Reference models/utils.py:87-178 is unreachable as written (IndexError at :136, index_select on
    a 0/1 mask at :140-142, exit() at :152). Its intent — class-agnostic NMS on each prior's best
    foreground score — has no runnable oracle, so parity is UNPINNED for this function. It is
    served here by the per-class path the working models use (detect)."""
    return detect(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy, config)
