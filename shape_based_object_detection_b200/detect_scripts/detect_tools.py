"""detect_scripts/detect_tools.py of the reference (detect, detect_refine, detect_objects) on the
fused CUDA eval path: softmax only, plus the class-agnostic second NMS at IoU 0.7 (:202-205,:324-327)."""
from ..core import detect_batched, unpad_detections

SECOND_NMS_IOU = 0.7


def detect(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy):
    """detect_tools.py:100-219: 'offset' decode, per-class NMS, second class-agnostic NMS(0.7); the
    result is in score order and is cut to top_k only when the first stage kept more than top_k."""
    out = detect_batched(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy,
                         act="softmax", box_type="offset", second_nms_thr=SECOND_NMS_IOU)
    return unpad_detections(out[0], out[1], out[2], out[4])


def detect_refine(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy,
                  prior_positives_idx=None):
    """detect_tools.py:222-341: predicted_locs are xyxy already and are clamped IN PLACE (:264)."""
    out = detect_batched(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy,
                         act="softmax", box_type="corner", clamp_inplace=True,
                         prior_keep=prior_positives_idx, second_nms_thr=SECOND_NMS_IOU)
    return unpad_detections(out[0], out[1], out[2], out[4])


def detect_objects(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy):
    """detect_tools.py:10-97 never returns (exit() at :69); parity UNPINNED. Intended semantics as
    models.utils.detect_objects: softmax, 'offset' decode, one candidate per prior (best foreground class),
    one class-agnostic NMS, label = arg-max class."""
    out = detect_batched(predicted_locs, predicted_scores, min_score, max_overlap, top_k, priors_cxcy,
                         act="softmax", box_type="offset", class_agnostic=True)
    return unpad_detections(out[0], out[1], out[2], out[4])
