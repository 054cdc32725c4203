"""models/utils.py:detect and detect_objects of the reference, on the fused CUDA eval path."""
import torch

from .. import _lib as L
from ..core import detect_batched, unpad_detections


def _cfg(config, key, default=None):
    try:
        return config[key]
    except (KeyError, TypeError, IndexError):
        try:
            return getattr(config, key)
        except (AttributeError, KeyError):
            return default


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
    """Reference models/utils.py:87-178 is unreachable as written (IndexError at :136, index_select on a
    0/1 mask at :140-142, exit() at :152), so parity is UNPINNED. This is its intended semantics ("each
    bounding box can only be assigned to one object", :92): every prior is ONE candidate scored with its best
    foreground probability (:135); candidates above min_score go through one class-agnostic NMS (:145); a kept
    box is labelled with its arg-max class (:147, read over the foreground columns: the class whose probability
    ordered the NMS); [0,0,1,1]/0/0.0 placeholder and top_k as in detect (:154-172)."""
    box_type = _cfg(config, "model")["box_type"]
    focal_type = _cfg(config, "focal_type")
    act = "sigmoid" if str(focal_type).lower() == "sigmoid" else "softmax"
    clamp_inplace = box_type not in ("offset", "center")
    out = detect_batched(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy,
                         act=act, box_type=box_type, clamp_inplace=clamp_inplace, class_agnostic=True)
    return unpad_detections(out[0], out[1], out[2], out[4])
