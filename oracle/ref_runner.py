"""CPU arm of bench.py (TEST / MEASUREMENT INFRASTRUCTURE ONLY — never imported by the product).

Runs the hot path on the host cores for one BASELINE config: the UNMODIFIED reference code when
oracle/_ref is present (oracle/build_ref.py put it there; `kind` = "reference"), otherwise the oracle port
(oracle/box_pipeline.py; `kind` = "port"). FCOS (config 5) is always the port: the reference's FCOSLoss does
not run (SURVEY.md §8 a-F).

Each runner exposes  train(locs..., boxes, labels) -> None (loss forward + backward)  and
eval(...) -> None (decode + threshold + NMS + top-k), on CPU tensors.
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


class Cfg(dict):
    """attr + item access like the reference's EasyDict configs."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _cfg(n_classes, reg="", cls="", box_type="offset"):
    return Cfg(device=torch.device("cpu"), n_classes=n_classes, reg_weights=1.0, reg_loss=reg, cls_loss=cls,
               model={"box_type": box_type}, focal_type="softmax")


def reference_available():
    from oracle import build_ref
    return build_ref.available()


_ref_modules = None


def _ref():
    """Import the reference's modules from oracle/_ref (once)."""
    global _ref_modules
    if _ref_modules is None:
        warnings.filterwarnings("ignore")
        saved = {k: sys.modules.get(k) for k in ("models", "operators", "metrics", "dataset", "detect_scripts")}
        sys.path.insert(0, REF)
        try:
            import importlib
            for k in saved:
                sys.modules.pop(k, None)
            mods = {
                "SSD300": importlib.import_module("models.SSD300"),
                "SSD512": importlib.import_module("models.SSD512"),
                "RetinaNet": importlib.import_module("models.RetinaNet"),
                "RefineDet512": importlib.import_module("models.RefineDet512"),
                "utils": importlib.import_module("models.utils"),
                "detect_tools": importlib.import_module("detect_scripts.detect_tools"),
                "Loss": importlib.import_module("operators.Loss"),
                "transforms": importlib.import_module("dataset.transforms"),
            }
        finally:
            sys.path.remove(REF)
        mods["detect_tools"].device = torch.device("cpu")  # module-global device of detect_tools.py:7
        _ref_modules = mods
    return _ref_modules


class Runner:
    """CPU implementation of one BASELINE config."""

    def __init__(self, config_id, priors_cxcy, n_classes, use_reference=None):
        self.id = config_id
        self.pri = priors_cxcy
        self.C = n_classes
        want_ref = reference_available() if use_reference is None else use_reference
        self.kind = "reference" if (want_ref and config_id != 5) else "port"
        self.crit = None
        if self.kind == "reference":
            R = _ref()
            if config_id == 1:
                self.crit = R["SSD300"].MultiBoxLoss300(self.pri, _cfg(n_classes))
            elif config_id == 2:
                self.crit = R["SSD512"].MultiBoxLoss512(self.pri, _cfg(n_classes))
            elif config_id == 3:
                # focal + GIoU: the reference class hard-wires IouLoss(losstype='Diou') on its 'DIOU' branch
                # (RetinaNet.py:369,461-463); the GIoU flavour is the reference's own operators.Loss.IouLoss
                self.crit = R["RetinaNet"].RetinaFocalLoss(self.pri, _cfg(n_classes, "DIOU", "FOCAL"))
                self.crit.Diou_loss = R["Loss"].IouLoss(pred_mode="Corner", reduce="mean", losstype="Giou")
            elif config_id == 4:
                self.crit = R["RefineDet512"].RefineDetLoss(self.pri, _cfg(n_classes))

    # ---- train: loss forward + backward ----
    def train(self, tensors, boxes, labels):
        from oracle import box_pipeline as O
        ts = [t.clone().requires_grad_(True) for t in tensors]
        if self.kind == "reference":
            loss = self.crit(*ts, boxes, labels)
        elif self.id == 1:
            loss = O.multibox_loss("s300", self.pri, ts[0], ts[1], boxes, labels)
        elif self.id == 2:
            loss = O.multibox_loss("s512", self.pri, ts[0], ts[1], boxes, labels)
        elif self.id == 3:
            loss = O.multibox_loss("ret", self.pri, ts[0], ts[1], boxes, labels, reg_loss="GIOU", cls_loss="FOCAL")
        elif self.id == 4:
            loss = O.refinedet_loss(self.pri, *ts, boxes, labels)
        else:
            loss = O.fcos_loss(self.pri, ts[0], ts[1], ts[2], boxes, labels, image_size=(800, 1333))
        loss.backward()
        return float(loss)

    # ---- eval: decode + threshold + NMS + top-k ----
    def eval(self, tensors, nms, keep=None):
        import torchvision
        from oracle import box_pipeline as O
        ms, mo, tk = nms
        if self.id in (1, 2, 3):
            locs, scores = tensors
            if self.kind == "reference":
                # (the reference has no per-class candidate cap: config 3's top-1000 is not applied on this arm)
                return _ref()["utils"].detect(locs.clone(), scores, ms, mo, tk, self.pri, _cfg(self.C))
            return O.detect(locs.clone(), scores, ms, mo, tk, self.pri, nms_fn=torchvision.ops.nms,
                            pre_nms_topk=1000 if self.id == 3 else 0)
        if self.id == 4:
            arm_l, odm_l, scores = tensors
            if self.kind == "reference":
                T = _ref()["transforms"]
                boxes = torch.stack([T.cxcy_to_xy(T.gcxgcy_to_cxcy(odm_l[i], T.gcxgcy_to_cxcy(arm_l[i], self.pri)))
                                     for i in range(arm_l.size(0))])  # RefineDet512.offset2bbox, :643-653
                return _ref()["detect_tools"].detect_refine(boxes, scores, ms, mo, tk, self.pri,
                                                            prior_positives_idx=keep)
            boxes = O.offset2bbox(arm_l, odm_l, self.pri)
            return O.detect(boxes, scores, ms, mo, tk, self.pri, box_type="corner", prior_keep=keep, second_nms=0.7,
                            nms_fn=torchvision.ops.nms)
        locs, scores, ctr = tensors
        bl, sc = O.fcos_postprocess(locs, scores, ctr, self.pri)
        return O.detect(bl, sc, ms, mo, tk, None, box_type="corner", focal_type="none_is_identity",
                        nms_fn=torchvision.ops.nms)
