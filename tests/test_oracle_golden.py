"""CPU suite: pin the oracle (oracle/box_pipeline.py) to the fixtures generated from the reference
itself (oracle/make_golden.py -> tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from cases import DETECT_CASES, LOSS_CASES, case_priors, operator_inputs
from oracle import box_pipeline as O
from shape_based_object_detection_b200 import priors as PR
from shape_based_object_detection_b200 import synth


def T(a):
    return torch.from_numpy(np.asarray(a))


def eq_nan(a, b):
    """bit-equality that treats NaN == NaN (CIoU of identical boxes is 0/0 in the reference)."""
    return torch.equal(a.isnan(), b.isnan()) and torch.equal(a.nan_to_num(7.0), b.nan_to_num(7.0))


def close_nan(a, b, rtol, atol):
    return torch.equal(a.isnan(), b.isnan()) and torch.allclose(a.nan_to_num(7.0), b.nan_to_num(7.0), rtol=rtol,
                                                                atol=atol)


def test_prior_tables_match_reference_generators(golden):
    meta = golden["priors_meta"]
    for k in ("ssd300", "ssd512", "retinanet", "refinedet512"):
        p = PR.PRIOR_TABLES[k]()
        assert p.shape[0] == int(meta[k + "_n"])
        assert abs(p.double().sum().item() - float(meta[k + "_sum"])) < 1e-9
    assert PR.ssd512_canonical_priors().shape[0] == 24564
    assert PR.retinanet640_priors().shape[0] == 76725


def test_dense_iou_bit_exact(golden):
    I, G = operator_inputs(), golden["operators"]
    assert torch.equal(O.find_jaccard_overlap(I["boxes"], I["sub"]), T(G["iou_metrics"]))
    ref = T(G["iou_jaccard"])
    got = O.jaccard(I["boxes"][1:], I["sub"])
    assert torch.equal(got.isnan(), ref.isnan())
    assert torch.equal(got.nan_to_num(7.0), ref.nan_to_num(7.0))
    assert torch.equal(O.intersect(I["boxes"], I["sub"]), T(G["intersect"]))


def test_converters_and_codec_bit_exact(golden):
    I, G = operator_inputs(), golden["operators"]
    assert torch.equal(O.xy_to_cxcy(I["sub"]), T(G["xy_to_cxcy"]))
    assert torch.equal(O.cxcy_to_xy(I["ppm"]), T(G["cxcy_to_xy"]))
    assert torch.equal(O.cxcy_to_gcxgcy(O.xy_to_cxcy(I["b1"]), I["ppm"]), T(G["enc_t"]))
    assert torch.equal(O.gcxgcy_to_cxcy(I["loc"], I["ppm"]), T(G["dec_t"]))
    assert torch.equal(O.encode(I["b1"], I["ppm"], [0.1, 0.2]), T(G["enc_u"]))
    assert torch.allclose(O.decode(I["loc"], I["ppm"], [0.1, 0.2]), T(G["dec_u"]), rtol=1e-6, atol=1e-7)
    assert torch.equal(O.offset2bbox(I["loc"][None], (I["loc"] * 0.5)[None], I["ppm"]), T(G["offset2bbox"]))


@pytest.mark.parametrize("kind", ["iou", "giou", "diou", "ciou"])
def test_pair_overlaps_and_grads(golden, kind):
    I, G = operator_inputs(), golden["operators"]
    x1, x2 = I["b1"].clone().requires_grad_(True), I["b2"].clone().requires_grad_(True)
    v = O.pair_overlap(x1, x2, kind)
    (v * I["wts"]).sum().backward()
    assert eq_nan(v.detach(), T(G["pair_" + kind]))
    assert close_nan(x1.grad, T(G["pair_" + kind + "_g1"]), 1e-6, 1e-6)
    assert close_nan(x2.grad, T(G["pair_" + kind + "_g2"]), 1e-6, 1e-6)


def test_row_losses(golden):
    I, G = operator_inputs(), golden["operators"]
    x = I["lg"].clone().requires_grad_(True)
    fl = O.focal_loss(x, I["tg"])
    fl.backward()
    assert abs(fl.item() - float(G["focal"])) <= 1e-6 * abs(float(G["focal"]))
    assert torch.allclose(x.grad, T(G["focal_g"]), rtol=1e-5, atol=1e-7)
    x = I["lg"].clone().requires_grad_(True)
    sf = O.sigmoid_focal_loss(x, I["tg"], 2.0, 0.25)
    sf.backward()
    assert abs(sf.item() - float(G["sigfocal"])) <= 1e-6 * abs(float(G["sigfocal"]))
    assert torch.allclose(x.grad, T(G["sigfocal_g"]), rtol=1e-5, atol=1e-7)
    x = I["pr"].clone().requires_grad_(True)
    s1 = O.smooth_l1_rows(x, I["tgt"]).sum() / x.size(0)
    s1.backward()
    assert abs(s1.item() - float(G["smoothl1"])) <= 1e-6 * abs(float(G["smoothl1"]))
    assert torch.allclose(x.grad, T(G["smoothl1_g"]), rtol=1e-6, atol=1e-8)
    for lt in ("Iou", "Giou", "Diou", "Ciou"):
        got = (1.0 - O.pair_overlap(I["b1"], I["b2"], lt.lower())).sum() / I["b1"].shape[0]
        ref = float(G["iouloss_" + lt])
        assert (np.isnan(ref) and np.isnan(got.item())) or abs(got.item() - ref) <= 1e-6


@pytest.mark.parametrize("thr", [0.5, 0.6])
def test_assignment_indices_bit_exact(golden, thr):
    I, G = operator_inputs(), golden["operators"]
    ov, obj, cls, neg = O.assign_image(I["boxes"], I["labels"], I["pri_xy"], thr)
    tag = "assign%02d_" % int(thr * 10)
    assert torch.equal(ov, T(G[tag + "ov"]))
    assert torch.equal(obj, T(G[tag + "obj"]).long())
    assert torch.equal(cls, T(G[tag + "cls"]).long())
    assert torch.equal(neg, T(G[tag + "neg"]).long())


def test_match_and_nms(golden):
    I, G = operator_inputs(), golden["operators"]
    for nm, enc in (("match", True), ("match_ious", False)):
        loc, conf = O.match(0.5, I["boxes"][1:], I["pri"], [0.1, 0.2], I["labels"][1:], encode_loc=enc)
        assert torch.equal(conf, T(G[nm + "_conf"]).long())
        assert torch.allclose(loc[conf > 0], T(G[nm + "_loc_pos"]), rtol=1e-6, atol=1e-7)
    keep = O.greedy_nms(I["nb"], I["ns"], 0.45)
    assert torch.equal(keep, T(G["tv_nms_keep"]).long())
    k200 = T(G["nms_keep"]).long()  # iou_utils.nms looks at the 200 best-scored boxes only
    assert torch.equal(keep[: k200.numel()], k200)
    import torchvision
    assert torch.equal(keep, torchvision.ops.nms(I["nb"], I["ns"], 0.45))


def _loss_inputs(case):
    pri = case_priors(case)
    locs, scores, bx, lb = synth.make_train_batch(pri, case["N"], case["C"], case["gmax"], case["seed"])
    if case.get("adversarial"):
        bx[0], lb[0] = synth.adversarial_gt(O.cxcy_to_xy(pri), case["C"])
    return pri, locs, scores, bx, lb


@pytest.mark.parametrize("name", [k for k, c in LOSS_CASES.items() if c["variant"] != "rfd"])
def test_loss_modules_match_reference(golden, name):
    case, G = LOSS_CASES[name], golden["losses"]
    pri, locs, scores, bx, lb = _loss_inputs(case)
    locs.requires_grad_(True)
    scores.requires_grad_(True)
    loss = O.multibox_loss(case["variant"], pri, locs, scores, bx, lb, reg_loss=case["reg"], cls_loss=case["cls"],
                           threshold=case.get("threshold", 0.5))
    loss.backward()
    ref = float(G[name + "_loss"])
    assert abs(loss.item() - ref) <= 1e-6 * abs(ref)
    assert torch.allclose(locs.grad, T(G[name + "_g_locs"]), rtol=1e-5, atol=1e-8)
    assert torch.allclose(scores.grad, T(G[name + "_g_scores"]), rtol=1e-5, atol=1e-8)


def refinedet_inputs(case):
    pri = case_priors(case)
    gen = torch.Generator().manual_seed(case["seed"])
    bx, lb = synth.make_gt(case["N"], case["gmax"], case["C"], gen, dense=True)
    P = pri.size(0)
    arm_l = torch.randn((case["N"], P, 4), generator=gen) * 0.1
    arm_s = torch.randn((case["N"], P, 2), generator=gen) * 2
    odm_l = torch.randn((case["N"], P, 4), generator=gen) * 0.1
    odm_s = torch.randn((case["N"], P, case["C"]), generator=gen)
    return pri, arm_l, arm_s, odm_l, odm_s, bx, lb


def test_refinedet_loss_matches_reference(golden):
    case, G = LOSS_CASES["rfd"], golden["losses"]
    pri, arm_l, arm_s, odm_l, odm_s, bx, lb = refinedet_inputs(case)
    ts = [t.requires_grad_(True) for t in (arm_l, arm_s, odm_l, odm_s)]
    loss = O.refinedet_loss(pri, *ts, bx, lb)
    loss.backward()
    ref = float(G["rfd_loss"])
    assert abs(loss.item() - ref) <= 1e-6 * abs(ref)
    for t, k in zip(ts, ("g_arm_l", "g_arm_s", "g_odm_l", "g_odm_s")):
        assert torch.allclose(t.grad, T(G["rfd_" + k]), rtol=1e-5, atol=1e-8)


def detect_inputs(case):
    pri = case_priors(case)
    locs, scores = synth.make_eval_batch(pri, case["N"], case["C"], case["seed"], bg_bias=case["bg"])
    keep = (scores[:, :, 1] > 0.0) if case.get("prior_keep") else None
    bt = case.get("box_type", "offset")
    if case["fn"] == "tools.refine":
        bt = "corner"
    if bt == "corner":
        locs = torch.stack([O.cxcy_to_xy(O.gcxgcy_to_cxcy(locs[i], pri)) for i in range(case["N"])])
    elif bt == "center":
        locs = torch.stack([O.gcxgcy_to_cxcy(locs[i], pri) for i in range(case["N"])])
    return pri, locs, scores, keep, bt


@pytest.mark.parametrize("name", list(DETECT_CASES))
def test_detect_matches_reference(golden, name):
    case, G = DETECT_CASES[name], golden["detect"]
    pri, locs, scores, keep, bt = detect_inputs(case)
    second = None if case["fn"] == "utils.detect" else 0.7
    b, l, s = O.detect(locs.clone(), scores, case["min_score"], case["max_overlap"], case["top_k"], pri,
                       box_type=bt, focal_type=case.get("focal_type", "softmax"), prior_keep=keep,
                       second_nms=second)
    for i in range(case["N"]):
        assert torch.equal(l[i], T(G[f"{name}_l{i}"]).long())
        assert torch.equal(s[i], T(G[f"{name}_s{i}"]))
        assert torch.equal(b[i], T(G[f"{name}_b{i}"]))


# ---------------------------------------------------------------------------------------------
# metrics.calculate_mAP
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["small", "voc_like", "strict", "sparse"])
def test_oracle_map_matches_reference(golden, name):
    from cases import MAP_CASES, map_inputs
    from oracle import box_pipeline as O
    case = MAP_CASES[name]
    ap, mean_ap = O.calculate_mAP(*map_inputs(case), case["threshold"], case["n_classes"])
    assert torch.equal(ap, T(golden["map"][name + "_ap"]))
    assert mean_ap == float(golden["map"][name + "_map"])


def test_oracle_bce_focal_matches_reference(golden):
    """Loss.py:83-103 FocalLoss (a12: defined by the reference, never called)."""
    I, G = operator_inputs(), golden["extras"]
    for tag, scale in (("", 1.0), ("_wide", 6.0)):
        x = (I["lg"] * scale).clone().requires_grad_(True)
        fl = O.bce_focal_loss(x, I["tg"], 0.25, 2)
        fl.backward()
        assert abs(fl.item() - float(G["bcefocal" + tag])) <= 1e-6 * abs(float(G["bcefocal" + tag]))
        assert torch.allclose(x.grad, T(G["bcefocal" + tag + "_g"]), rtol=1e-5, atol=1e-7)


def test_oracle_diounms_matches_reference(golden):
    """iou_utils.diounms (a20: defined by the reference, never called)."""
    I, G = operator_inputs(), golden["extras"]
    for beta in (1.0, 0.6):
        keep, cnt = O.diounms(I["nb"], I["ns"], 0.45, 200, beta)
        assert torch.equal(keep[:cnt], T(G["diounms_b%02d" % int(beta * 10)]).long())
