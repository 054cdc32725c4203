"""Switch the reference over to the CUDA path without editing it."""
import importlib
import sys


def install(reference_root=None):
    """Import the reference's modules (from reference_root if given) and rebind the hot-path names to
    this package's implementations. Returns the list of patched attributes."""
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from . import metrics as m
    from .dataset import transforms as t
    from .detect_scripts import detect_tools as dt
    from .models import FCOSDet, RefineDet512, RetinaNet, SSD300, SSD512, utils as mu
    from .operators import Loss as lo, iou_utils as iu

    plan = [
        ("operators.iou_utils", iu, ["bbox_overlaps_iou", "bbox_overlaps_giou", "bbox_overlaps_diou",
                                     "bbox_overlaps_ciou", "point_form", "center_size", "intersect", "jaccard",
                                     "match", "match_ious", "encode", "decode", "nms", "diounms"]),
        ("operators.Loss", lo, ["focal_loss", "SigmoidFocalLoss", "FocalLoss", "IouLoss", "SmoothL1Loss"]),
        ("models.SSD300", SSD300, ["MultiBoxLoss300"]),
        ("models.SSD512", SSD512, ["MultiBoxLoss512"]),
        ("models.RetinaNet", RetinaNet, ["RetinaFocalLoss"]),
        ("models.RefineDet512", RefineDet512, ["RefineDetLoss"]),
        ("models.FCOSDet", FCOSDet, ["FCOSLoss"]),
        ("models.utils", mu, ["detect", "detect_objects"]),
        ("models", None, []),
        ("detect_scripts.detect_tools", dt, ["detect", "detect_refine", "detect_objects"]),
    ]
    patched = []
    for mod_name, ours, names in plan:
        try:
            ref = importlib.import_module(mod_name)
        except Exception:  # the reference module may need packages that are not installed
            continue
        for name in names:
            setattr(ref, name, getattr(ours, name))
            patched.append(f"{mod_name}.{name}")
    # models/__init__.py re-exports the loss classes it imported at load time (models/__init__.py:1-5)
    try:
        ref_models = importlib.import_module("models")
        for cls_name, ours in (("MultiBoxLoss300", SSD300), ("MultiBoxLoss512", SSD512),
                               ("RetinaFocalLoss", RetinaNet), ("RefineDetLoss", RefineDet512),
                               ("FCOSLoss", FCOSDet)):
            if hasattr(ref_models, cls_name):
                setattr(ref_models, cls_name, getattr(ours, cls_name))
                patched.append(f"models.{cls_name}")
    except Exception:
        pass
    # metrics.find_jaccard_overlap is also called on CPU tensors inside DataLoader workers
    # (dataset/transforms.py:175): leave metrics.py alone and patch only the model-side imports.
    for mod_name in ("models.SSD300", "models.SSD512", "models.RetinaNet", "models.RefineDet512"):
        try:
            ref = importlib.import_module(mod_name)
            if hasattr(ref, "find_jaccard_overlap"):
                setattr(ref, "find_jaccard_overlap", m.find_jaccard_overlap)
                patched.append(f"{mod_name}.find_jaccard_overlap")
        except Exception:
            pass
    # metrics.calculate_mAP (train_*.py / eval.py import it by name after install): only this
    # function of metrics.py is replaced
    try:
        ref_metrics = importlib.import_module("metrics")
        if hasattr(ref_metrics, "calculate_mAP"):
            setattr(ref_metrics, "calculate_mAP", m.calculate_mAP)
            patched.append("metrics.calculate_mAP")
    except Exception:
        pass
    return patched
