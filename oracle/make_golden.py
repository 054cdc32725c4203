"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU (build container only).

Usage:  python oracle/make_golden.py [/root/reference]
The reference cannot travel to the GPU box, so its outputs on seeded synthetic inputs are committed
as small fixtures; inputs are regenerated from the same seeds by the tests
(shape_based_object_detection_b200.synth / .priors), never stored.
"""
import os
import sys
import warnings

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_args = [a for a in sys.argv[1:] if not a.startswith("--")]
REF = _args[0] if _args else "/root/reference"
ONLY_MAP = "--only-map" in sys.argv  # regenerate tests/golden/map.npz and nothing else
ONLY_EXTRAS = "--only-extras" in sys.argv  # regenerate tests/golden/extras.npz and nothing else
ONLY_CROP = "--only-crop" in sys.argv  # regenerate tests/golden/crop.npz and nothing else
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from shape_based_object_detection_b200 import priors as PR  # noqa: E402
from shape_based_object_detection_b200 import synth  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


class Cfg(dict):
    """attr + item access like EasyDict (absent in this image)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def cfg(reg="", cls="", n_classes=6, box_type="offset", focal_type="softmax"):
    return Cfg(device=torch.device("cpu"), n_classes=n_classes, reg_weights=1.0, reg_loss=reg, cls_loss=cls,
               model={"box_type": box_type}, focal_type=focal_type, nms={"min_score": 0.01, "max_overlap": 0.45,
                                                                         "top_k": 200})


class Stub:
    device = torch.device("cpu")


def golden_map():
    """metrics.calculate_mAP (metrics.py:8-145) on CPU for tests/cases.py:MAP_CASES."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from cases import MAP_CASES, map_inputs  # noqa: E402
    import metrics as ref_metrics
    out = {}
    for name, case in MAP_CASES.items():
        label_map = {("background" if i == 0 else "c%d" % i): i for i in range(case["n_classes"])}
        aps, mean_ap = ref_metrics.calculate_mAP(*map_inputs(case), case["threshold"], label_map, device="cpu")
        out[name + "_ap"] = np.array([aps["c%d" % i] for i in range(1, case["n_classes"])], dtype=np.float32)
        out[name + "_map"] = np.array(mean_ap, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "map.npz"), **out)
    print("map ok", {k: float(v) for k, v in out.items() if k.endswith("_map")})


def golden_crop():
    """dataset/transforms.py:124-205 random_crop on CPU: outputs and the NEXT draw of python's random (the state
    the function leaves behind) for 24 seeds."""
    import random
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from cases import crop_inputs  # noqa: E402
    from dataset import transforms as ref_t
    out = {}
    for seed in range(24):
        image, boxes, labels = crop_inputs(seed)
        random.seed(1000 + seed)
        ni, nb, nl = ref_t.random_crop(image, boxes.clone(), labels.clone())
        out["s%d_shape" % seed] = np.array(ni.shape, dtype=np.int64)
        out["s%d_sum" % seed] = np.float64(ni.double().sum().item())
        out["s%d_boxes" % seed] = nb.numpy()
        out["s%d_labels" % seed] = nl.numpy()
        out["s%d_next" % seed] = np.float64(random.random())
    np.savez_compressed(os.path.join(OUT, "crop.npz"), **out)
    print("crop ok", sum(1 for s in range(24) if tuple(out["s%d_shape" % s]) != tuple(crop_inputs(s)[0].shape)), "of 24 cropped")


def golden_extras():
    """Functions of the path that the reference defines but never calls."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from cases import operator_inputs  # noqa: E402
    from operators import Loss as ref_loss
    I = operator_inputs()
    out = {}
    for tag, lg in (("", I["lg"]), ("_wide", I["lg"] * 6.0)):  # the wide logits reach the probability clamp
        x = lg.clone().requires_grad_(True)
        fl = ref_loss.FocalLoss(0.25, 2)(x, I["tg"])
        fl.backward()
        out["bcefocal" + tag] = np.float64(fl.item())
        out["bcefocal" + tag + "_g"] = x.grad.numpy()
    from operators import iou_utils as ref_iu
    for beta in (1.0, 0.6):
        keep, cnt = ref_iu.diounms(I["nb"].clone(), I["ns"].clone(), 0.45, 200, beta)
        out["diounms_b%02d" % int(beta * 10)] = keep[:cnt].numpy().astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "extras.npz"), **out)
    print("extras ok", {k: (float(v) if v.ndim == 0 else v.shape) for k, v in out.items() if not k.endswith("_g")})


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    if ONLY_MAP:
        golden_map()
        return
    if ONLY_EXTRAS:
        golden_extras()
        return
    if ONLY_CROP:
        golden_crop()
        return
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from cases import LOSS_CASES, DETECT_CASES, case_priors  # noqa: E402

    import metrics as ref_metrics
    from dataset import transforms as ref_t
    from detect_scripts import detect_tools as ref_dt
    from models import utils as ref_mu
    from models.RefineDet512 import RefineDet512, RefineDetLoss
    from models.RetinaNet import RetinaFocalLoss, RetinaNet
    from models.SSD300 import SSD300, MultiBoxLoss300
    from models.SSD512 import SSD512, MultiBoxLoss512
    from operators import Loss as ref_loss
    from operators import iou_utils as ref_iu

    # ---- prior tables == the reference generators ------------------------------------------
    checks = {
        "ssd300": SSD300.create_prior_boxes(Stub()), "ssd512": SSD512.create_prior_boxes(Stub()),
        "retinanet": RetinaNet.create_anchors(Stub()), "refinedet512": RefineDet512.create_prior_boxes(Stub()),
    }
    pri_meta = {}
    for k, ref in checks.items():
        mine = PR.PRIOR_TABLES[k]()
        assert mine.shape == ref.shape and torch.equal(mine, ref), k
        pri_meta[k + "_n"] = np.int64(ref.shape[0])
        pri_meta[k + "_sum"] = np.float64(ref.double().sum().item())
    np.savez_compressed(os.path.join(OUT, "priors_meta.npz"), **pri_meta)
    print("priors ok", {k: int(v) for k, v in pri_meta.items() if k.endswith("_n")})

    # ---- operators: IoU matrix, converters, codec, paired IoU, row losses, match, nms -------
    from cases import operator_inputs  # noqa: E402
    I = operator_inputs()
    pri, boxes, labels, sub, b1, b2, ppm, loc = (I["pri"], I["boxes"], I["labels"], I["sub"], I["b1"], I["b2"],
                                                 I["ppm"], I["loc"])
    assert torch.equal(I["pri_xy"], ref_t.cxcy_to_xy(pri))
    ops = {}
    ops["iou_metrics"] = ref_metrics.find_jaccard_overlap(boxes, sub).numpy()
    ops["iou_jaccard"] = ref_iu.jaccard(boxes[1:], sub).numpy()
    ops["intersect"] = ref_iu.intersect(boxes, sub).numpy()
    ops["xy_to_cxcy"] = ref_t.xy_to_cxcy(sub).numpy()
    ops["cxcy_to_xy"] = ref_t.cxcy_to_xy(ppm).numpy()
    for kind, fn in (("iou", ref_iu.bbox_overlaps_iou), ("giou", ref_iu.bbox_overlaps_giou),
                     ("diou", ref_iu.bbox_overlaps_diou), ("ciou", ref_iu.bbox_overlaps_ciou)):
        x1 = b1.clone().requires_grad_(True)
        x2 = b2.clone().requires_grad_(True)
        v = fn(x1, x2)
        (v * I["wts"]).sum().backward()
        ops["pair_" + kind] = v.detach().numpy()
        ops["pair_" + kind + "_g1"] = x1.grad.numpy()
        ops["pair_" + kind + "_g2"] = x2.grad.numpy()
    ops["enc_t"] = ref_t.cxcy_to_gcxgcy(ref_t.xy_to_cxcy(b1), ppm).numpy()
    ops["dec_t"] = ref_t.gcxgcy_to_cxcy(loc, ppm).numpy()
    ops["enc_u"] = ref_iu.encode(b1, ppm, [0.1, 0.2]).numpy()
    ops["dec_u"] = ref_iu.decode(loc, ppm, [0.1, 0.2]).numpy()
    ops["offset2bbox"] = RefineDet512.offset2bbox(
        type("S", (), {"priors_cxcy": ppm, "device": torch.device("cpu")})(), loc[None], (loc * 0.5)[None]).numpy()
    # row losses
    lg, tg = I["lg"], I["tg"]
    x = lg.clone().requires_grad_(True)
    fl = ref_loss.focal_loss(x, tg, device="cpu")
    fl.backward()
    ops["focal"] = np.float64(fl.item())
    ops["focal_g"] = x.grad.numpy()
    x = lg.clone().requires_grad_(True)
    sf = ref_loss.SigmoidFocalLoss(2.0, 0.25, cfg())(x, tg)
    sf.backward()
    ops["sigfocal"] = np.float64(sf.item())
    ops["sigfocal_g"] = x.grad.numpy()
    x = I["pr"].clone().requires_grad_(True)
    s1 = ref_loss.SmoothL1Loss()(x, I["tgt"])
    s1.backward()
    ops["smoothl1"] = np.float64(s1.item())
    ops["smoothl1_g"] = x.grad.numpy()
    for lt in ("Iou", "Giou", "Diou", "Ciou"):
        ops["iouloss_" + lt] = np.float64(ref_loss.IouLoss(losstype=lt)(b1, b2).item())
    # match / match_ious on the adversarial image (index 0 is degenerate: 0/0 -> NaN rows, skip it)
    P = pri.size(0)
    for nm, fn in (("match", ref_iu.match), ("match_ious", ref_iu.match_ious)):
        loc_t = torch.zeros((1, P, 4))
        conf_t = torch.zeros((1, P), dtype=torch.long)
        fn(0.5, boxes[1:], pri, [0.1, 0.2], labels[1:], loc_t, conf_t, 0)
        ops[nm + "_conf"] = conf_t[0].numpy().astype(np.int16)
        posm = conf_t[0] > 0
        ops[nm + "_loc_pos"] = loc_t[0][posm].numpy()
    # python nms of iou_utils / torchvision nms on random boxes
    nb, ns = I["nb"], I["ns"]
    keep, cnt = ref_iu.nms(nb, ns, 0.45, 200)
    ops["nms_keep"] = keep[:cnt].numpy().astype(np.int32)
    import torchvision
    ops["tv_nms_keep"] = torchvision.ops.nms(nb, ns, 0.45).numpy().astype(np.int32)
    # assignment intermediates via the reference primitives, in the reference's order (SSD512.py:535-563)
    for thr in (0.5, 0.6):
        overlap = ref_metrics.find_jaccard_overlap(boxes, I["pri_xy"])
        ov, obj = overlap.max(dim=0)
        ovo, pfo = overlap.max(dim=1)
        pfo = pfo[ovo > 0]
        if len(pfo) > 0:
            ov.index_fill_(0, pfo, 1.0)
        for j in range(pfo.size(0)):
            obj[pfo[j]] = j
        lab = labels[obj]
        neg = labels[obj]
        lab[ov < thr] = 0
        neg[ov < thr - 0.1] = -1
        tag = "assign%02d_" % int(thr * 10)
        ops[tag + "ov"] = ov.numpy()
        ops[tag + "obj"] = obj.numpy().astype(np.int16)
        ops[tag + "cls"] = lab.numpy().astype(np.int16)
        ops[tag + "neg"] = neg.numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **ops)
    print("operators ok")

    # ---- loss modules ----------------------------------------------------------------------
    ref_cls = {"s300": MultiBoxLoss300, "s512": MultiBoxLoss512, "ret": RetinaFocalLoss}
    out = {}
    for name, case in LOSS_CASES.items():
        pri = case_priors(case)
        if case["variant"] == "rfd":
            gen = torch.Generator().manual_seed(case["seed"])
            bx, lb = synth.make_gt(case["N"], case["gmax"], case["C"], gen, dense=True)
            Pn = pri.size(0)
            arm_l = (torch.randn((case["N"], Pn, 4), generator=gen) * 0.1).requires_grad_(True)
            arm_s = (torch.randn((case["N"], Pn, 2), generator=gen) * 2).requires_grad_(True)
            odm_l = (torch.randn((case["N"], Pn, 4), generator=gen) * 0.1).requires_grad_(True)
            odm_s = torch.randn((case["N"], Pn, case["C"]), generator=gen).requires_grad_(True)
            crit = RefineDetLoss(pri, cfg(n_classes=case["C"]))
            loss = crit(arm_l, arm_s, odm_l, odm_s, bx, lb)
            loss.backward()
            out[name + "_loss"] = np.float64(loss.item())
            out[name + "_g_arm_l"] = arm_l.grad.numpy()
            out[name + "_g_arm_s"] = arm_s.grad.numpy()
            out[name + "_g_odm_l"] = odm_l.grad.numpy()
            out[name + "_g_odm_s"] = odm_s.grad.numpy()
            continue
        locs, scores, bx, lb = synth.make_train_batch(pri, case["N"], case["C"], case["gmax"], case["seed"])
        if case.get("adversarial"):
            ab, al = synth.adversarial_gt(ref_t.cxcy_to_xy(pri), case["C"])
            bx[0], lb[0] = ab, al
        locs.requires_grad_(True)
        scores.requires_grad_(True)
        crit = ref_cls[case["variant"]](pri, cfg(case["reg"], case["cls"], case["C"]),
                                        threshold=case.get("threshold", 0.5))
        loss = crit(locs, scores, bx, lb)
        loss.backward()
        out[name + "_loss"] = np.float64(loss.item())
        out[name + "_g_locs"] = locs.grad.numpy()
        out[name + "_g_scores"] = scores.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)
    print("losses ok", {k: float(v) for k, v in out.items() if k.endswith("_loss")})

    # ---- eval path -----------------------------------------------------------------------
    dout = {}
    for name, case in DETECT_CASES.items():
        pri = case_priors(case)
        locs, scores = synth.make_eval_batch(pri, case["N"], case["C"], case["seed"], bg_bias=case["bg"])
        keep = None
        if case.get("prior_keep"):
            keep = scores[:, :, 1] > 0.0
        fn = case["fn"]
        if fn == "utils.detect":
            c = cfg(n_classes=case["C"], box_type=case.get("box_type", "offset"),
                    focal_type=case.get("focal_type", "softmax"))
            l_in = locs.clone()
            if case.get("box_type") == "corner":
                l_in = torch.stack([ref_t.cxcy_to_xy(ref_t.gcxgcy_to_cxcy(locs[i], pri)) for i in range(case["N"])])
            elif case.get("box_type") == "center":
                l_in = torch.stack([ref_t.gcxgcy_to_cxcy(locs[i], pri) for i in range(case["N"])])
            b, l, s = ref_mu.detect(l_in, scores, case["min_score"], case["max_overlap"], case["top_k"], pri, c,
                                    prior_positives_idx=keep)
        elif fn == "tools.detect":
            b, l, s = ref_dt.detect(locs.clone(), scores, case["min_score"], case["max_overlap"], case["top_k"], pri)
        else:
            l_in = torch.stack([ref_t.cxcy_to_xy(ref_t.gcxgcy_to_cxcy(locs[i], pri)) for i in range(case["N"])])
            b, l, s = ref_dt.detect_refine(l_in, scores, case["min_score"], case["max_overlap"], case["top_k"],
                                           pri, prior_positives_idx=keep)
        for i in range(case["N"]):
            dout[f"{name}_b{i}"] = b[i].numpy()
            dout[f"{name}_l{i}"] = l[i].numpy().astype(np.int16)
            dout[f"{name}_s{i}"] = s[i].numpy()
    np.savez_compressed(os.path.join(OUT, "detect.npz"), **dout)
    print("detect ok", {k: v.shape for k, v in dout.items() if "_s" in k})
    golden_map()
    golden_extras()
    golden_crop()


if __name__ == "__main__":
    main()
