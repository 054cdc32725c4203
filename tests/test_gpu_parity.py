"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the
fixtures generated from the reference. Tolerances (BASELINE.json north_star): indices / labels /
kept boxes bit-exact; encoded targets, losses, decoded boxes within 1e-5 relative; gradients within
1e-4 relative."""
import numpy as np
import pytest
import torch

from cases import DETECT_CASES, LOSS_CASES, case_priors, operator_inputs
from test_oracle_golden import T, _loss_inputs, close_nan, detect_inputs, eq_nan, refinedet_inputs

pytestmark = pytest.mark.gpu

RTOL_LOSS = 1e-5
RTOL_GRAD = 1e-4


class Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def cfg(reg="", cls="", n_classes=6, box_type="offset", focal_type="softmax"):
    return Cfg(device=torch.device("cuda:0"), n_classes=n_classes, reg_weights=1.0, reg_loss=reg, cls_loss=cls,
               model={"box_type": box_type}, focal_type=focal_type)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "the gpu suite needs a CUDA device"
    return torch.device("cuda:0")


def cu(x, dev):
    return x.to(dev)


# ---------------------------------------------------------------------------------------------
# stand-alone operators
# ---------------------------------------------------------------------------------------------
def test_dense_iou_bit_exact(golden, dev):
    from shape_based_object_detection_b200.metrics import find_jaccard_overlap
    from shape_based_object_detection_b200.operators import iou_utils as U
    I, G = operator_inputs(), golden["operators"]
    got = find_jaccard_overlap(cu(I["boxes"], dev), cu(I["sub"], dev)).cpu()
    assert torch.equal(got, T(G["iou_metrics"]))
    assert eq_nan(U.jaccard(cu(I["boxes"][1:], dev), cu(I["sub"], dev)).cpu(), T(G["iou_jaccard"]))
    assert torch.equal(U.intersect(cu(I["boxes"], dev), cu(I["sub"], dev)).cpu(), T(G["intersect"]))
    # full-size: every SSD300 prior
    from oracle import box_pipeline as O
    want = O.find_jaccard_overlap(I["boxes"], I["pri_xy"])
    assert torch.equal(find_jaccard_overlap(cu(I["boxes"], dev), cu(I["pri_xy"], dev)).cpu(), want)
    # empty inputs
    assert find_jaccard_overlap(torch.zeros((0, 4), device=dev), cu(I["sub"], dev)).shape == (0, I["sub"].size(0))


def test_converters_and_codec(golden, dev):
    from shape_based_object_detection_b200.dataset import transforms as TR
    from shape_based_object_detection_b200.models.RefineDet512 import offset2bbox
    from shape_based_object_detection_b200.operators import iou_utils as U
    I, G = operator_inputs(), golden["operators"]
    assert torch.equal(TR.xy_to_cxcy(cu(I["sub"], dev)).cpu(), T(G["xy_to_cxcy"]))
    assert torch.equal(TR.cxcy_to_xy(cu(I["ppm"], dev)).cpu(), T(G["cxcy_to_xy"]))
    assert torch.equal(U.point_form(cu(I["ppm"], dev)).cpu(), T(G["cxcy_to_xy"]))
    ppm = cu(I["ppm"], dev)
    enc = TR.cxcy_to_gcxgcy(TR.xy_to_cxcy(cu(I["b1"], dev)), ppm).cpu()
    assert torch.allclose(enc, T(G["enc_t"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(TR.gcxgcy_to_cxcy(cu(I["loc"], dev), ppm).cpu(), T(G["dec_t"]), rtol=1e-5, atol=1e-7)
    assert torch.allclose(U.encode(cu(I["b1"], dev), ppm, [0.1, 0.2]).cpu(), T(G["enc_u"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(U.decode(cu(I["loc"], dev), ppm, [0.1, 0.2]).cpu(), T(G["dec_u"]), rtol=1e-5, atol=1e-7)
    o2b = offset2bbox(cu(I["loc"][None], dev), cu((I["loc"] * 0.5)[None], dev), ppm).cpu()
    assert torch.allclose(o2b, T(G["offset2bbox"]), rtol=1e-5, atol=1e-7)
    # round trip property at full size
    pri = cu(I["pri"], dev)
    rt = TR.xy_to_cxcy(TR.cxcy_to_xy(pri))
    assert torch.allclose(rt, pri, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("kind", ["iou", "giou", "diou", "ciou"])
def test_pair_overlaps_and_grads(golden, dev, kind):
    from shape_based_object_detection_b200.operators import iou_utils as U
    I, G = operator_inputs(), golden["operators"]
    fn = getattr(U, "bbox_overlaps_" + kind)
    x1 = cu(I["b1"], dev).requires_grad_(True)
    x2 = cu(I["b2"], dev).requires_grad_(True)
    v = fn(x1, x2)
    (v * cu(I["wts"], dev)).sum().backward()
    assert close_nan(v.detach().cpu(), T(G["pair_" + kind]), 1e-5, 1e-6)
    assert close_nan(x1.grad.cpu(), T(G["pair_" + kind + "_g1"]), RTOL_GRAD, 1e-4)
    assert close_nan(x2.grad.cpu(), T(G["pair_" + kind + "_g2"]), RTOL_GRAD, 1e-4)


def test_row_losses(golden, dev):
    from shape_based_object_detection_b200.operators import Loss as LS
    I, G = operator_inputs(), golden["operators"]
    x = cu(I["lg"], dev).requires_grad_(True)
    fl = LS.focal_loss(x, cu(I["tg"], dev), device=dev)
    fl.backward()
    assert abs(fl.item() - float(G["focal"])) <= RTOL_LOSS * abs(float(G["focal"]))
    assert torch.allclose(x.grad.cpu(), T(G["focal_g"]), rtol=RTOL_GRAD, atol=1e-6)
    x = cu(I["lg"], dev).requires_grad_(True)
    sf = LS.SigmoidFocalLoss(2.0, 0.25, cfg())(x, cu(I["tg"], dev))
    sf.backward()
    assert abs(sf.item() - float(G["sigfocal"])) <= RTOL_LOSS * abs(float(G["sigfocal"]))
    assert torch.allclose(x.grad.cpu(), T(G["sigfocal_g"]), rtol=RTOL_GRAD, atol=1e-6)
    x = cu(I["pr"], dev).requires_grad_(True)
    s1 = LS.SmoothL1Loss()(x, cu(I["tgt"], dev))
    s1.backward()
    assert abs(s1.item() - float(G["smoothl1"])) <= RTOL_LOSS * abs(float(G["smoothl1"]))
    assert torch.allclose(x.grad.cpu(), T(G["smoothl1_g"]), rtol=RTOL_GRAD, atol=1e-8)
    for lt in ("Iou", "Giou", "Diou"):
        got = LS.IouLoss(losstype=lt)(cu(I["b1"], dev), cu(I["b2"], dev)).item()
        assert abs(got - float(G["iouloss_" + lt])) <= 1e-5


@pytest.mark.parametrize("thr", [0.5, 0.6])
def test_assignment_indices_bit_exact(golden, dev, thr):
    import shape_based_object_detection_b200 as S
    I, G = operator_inputs(), golden["operators"]
    ov, obj, cls, neg = S.assign([cu(I["boxes"], dev)], [cu(I["labels"], dev)], cu(I["pri_xy"], dev), threshold=thr)
    tag = "assign%02d_" % int(thr * 10)
    assert torch.equal(ov[0].cpu(), T(G[tag + "ov"]))
    assert torch.equal(obj[0].cpu().long(), T(G[tag + "obj"]).long())
    assert torch.equal(cls[0].cpu(), T(G[tag + "cls"]).long())
    assert torch.equal(neg[0].cpu(), T(G[tag + "neg"]).long())


def test_match_and_nms(golden, dev):
    from shape_based_object_detection_b200.operators import iou_utils as U
    I, G = operator_inputs(), golden["operators"]
    P = I["pri"].size(0)
    for nm, fn in (("match", U.match), ("match_ious", U.match_ious)):
        loc_t = torch.zeros((1, P, 4), device=dev)
        conf_t = torch.zeros((1, P), dtype=torch.long, device=dev)
        fn(0.5, cu(I["boxes"][1:], dev), cu(I["pri"], dev), [0.1, 0.2], cu(I["labels"][1:], dev), loc_t, conf_t, 0)
        conf = conf_t[0].cpu()
        assert torch.equal(conf, T(G[nm + "_conf"]).long())
        assert torch.allclose(loc_t[0].cpu()[conf > 0], T(G[nm + "_loc_pos"]), rtol=1e-5, atol=1e-6)
    keep = U.torchvision_nms(cu(I["nb"], dev), cu(I["ns"], dev), 0.45).cpu()
    assert torch.equal(keep, T(G["tv_nms_keep"]).long())
    k, cnt = U.nms(cu(I["nb"], dev), cu(I["ns"], dev), 0.45, 200)
    assert torch.equal(k[:cnt].cpu(), T(G["nms_keep"]).long())
    assert k.numel() == I["nb"].size(0) and int(k[cnt:].abs().sum()) == 0
    # bigger random problem against the oracle's greedy NMS
    from oracle import box_pipeline as O
    g = torch.Generator().manual_seed(5)
    c = torch.rand((3000, 2), generator=g)
    wh = torch.rand((3000, 2), generator=g) * 0.2 + 0.01
    b = torch.cat([c - wh / 2, c + wh / 2], 1)
    s = torch.rand((3000,), generator=g)
    s[100:110] = s[100]  # score ties -> lower index first
    assert torch.equal(U.torchvision_nms(cu(b, dev), cu(s, dev), 0.3).cpu(), O.greedy_nms(b, s, 0.3))


# ---------------------------------------------------------------------------------------------
# loss modules
# ---------------------------------------------------------------------------------------------
def _module(variant):
    from shape_based_object_detection_b200 import models as M
    return {"s300": M.MultiBoxLoss300, "s512": M.MultiBoxLoss512, "ret": M.RetinaFocalLoss}[variant]


@pytest.mark.parametrize("name", [k for k, c in LOSS_CASES.items() if c["variant"] != "rfd"])
def test_loss_modules(golden, dev, name):
    from oracle import box_pipeline as O
    case, G = LOSS_CASES[name], golden["losses"]
    pri, locs, scores, bx, lb = _loss_inputs(case)
    crit = _module(case["variant"])(cu(pri, dev), cfg(case["reg"], case["cls"], case["C"]),
                                    threshold=case.get("threshold", 0.5))
    l_d = cu(locs, dev).requires_grad_(True)
    s_d = cu(scores, dev).requires_grad_(True)
    loss = crit(l_d, s_d, [cu(b, dev) for b in bx], [cu(l, dev) for l in lb])
    loss.backward()
    ref = float(G[name + "_loss"])
    assert abs(loss.item() - ref) <= RTOL_LOSS * abs(ref), (loss.item(), ref)
    # intermediates against the oracle (itself pinned to the reference by the CPU suite)
    _, parts = O.multibox_loss(case["variant"], pri, locs, scores, bx, lb, reg_loss=case["reg"],
                               cls_loss=case["cls"], threshold=case.get("threshold", 0.5), want_parts=True)
    st = crit.last["state"]
    cls, neg = st.targets()
    assert torch.equal(st.obj.cpu().long(), parts["obj"])
    assert torch.equal(st.ov.cpu(), parts["ov"])
    assert torch.equal(cls.cpu(), parts["true_classes"])
    assert torch.equal(neg.cpu(), parts["true_neg_classes"])
    assert int(st.loss[3].item()) == int(parts["n_pos"].sum())
    assert torch.allclose(l_d.grad.cpu(), T(G[name + "_g_locs"]), rtol=RTOL_GRAD, atol=1e-7)
    assert torch.allclose(s_d.grad.cpu(), T(G[name + "_g_scores"]), rtol=RTOL_GRAD, atol=1e-7)


def test_refinedet_loss(golden, dev):
    from shape_based_object_detection_b200.models import RefineDetLoss
    case, G = LOSS_CASES["rfd"], golden["losses"]
    pri, arm_l, arm_s, odm_l, odm_s, bx, lb = refinedet_inputs(case)
    crit = RefineDetLoss(cu(pri, dev), cfg(n_classes=case["C"]))
    ts = [cu(t, dev).requires_grad_(True) for t in (arm_l, arm_s, odm_l, odm_s)]
    loss = crit(*ts, [cu(b, dev) for b in bx], [cu(l, dev) for l in lb])
    loss.backward()
    ref = float(G["rfd_loss"])
    assert abs(loss.item() - ref) <= RTOL_LOSS * abs(ref), (loss.item(), ref)
    for t, k in zip(ts, ("g_arm_l", "g_arm_s", "g_odm_l", "g_odm_s")):
        want = T(G["rfd_" + k])
        if t.grad is None:
            assert float(want.abs().max()) == 0.0
        else:
            assert torch.allclose(t.grad.cpu(), want, rtol=RTOL_GRAD, atol=1e-7), k


def test_loss_at_ssd512_coco_shape(dev):
    """P = 24 564, C = 81, G <= 100 (BASELINE config 2), N small enough for the CPU oracle."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512, RetinaFocalLoss
    pri = PR.ssd512_canonical_priors()
    locs, scores, bx, lb = synth.make_train_batch(pri, 3, 81, 100, 1234 + 2)
    for Mod, variant, reg, cls in ((MultiBoxLoss512, "s512", "", ""), (RetinaFocalLoss, "ret", "DIOU", "FOCAL")):
        l_c = locs.clone().requires_grad_(True)
        s_c = scores.clone().requires_grad_(True)
        want, parts = O.multibox_loss(variant, pri, l_c, s_c, bx, lb, reg_loss=reg, cls_loss=cls, want_parts=True)
        want.backward()
        crit = Mod(cu(pri, dev), cfg(reg, cls, 81))
        l_d = cu(locs, dev).requires_grad_(True)
        s_d = cu(scores, dev).requires_grad_(True)
        loss = crit(l_d, s_d, [cu(b, dev) for b in bx], [cu(l, dev) for l in lb])
        loss.backward()
        st = crit.last["state"]
        cls_t, neg_t = st.targets()
        assert torch.equal(st.obj.cpu().long(), parts["obj"])
        assert torch.equal(cls_t.cpu(), parts["true_classes"])
        assert torch.equal(neg_t.cpu(), parts["true_neg_classes"])
        assert abs(loss.item() - want.item()) <= RTOL_LOSS * abs(want.item()), (loss.item(), want.item())
        assert torch.allclose(l_d.grad.cpu(), l_c.grad, rtol=RTOL_GRAD, atol=1e-7)
        assert torch.allclose(s_d.grad.cpu(), s_c.grad, rtol=RTOL_GRAD, atol=1e-7)


def test_loss_full_batch_properties(dev):
    """BASELINE size (N=32, P=24 564, C=81): size-independent properties instead of the oracle —
    the batch sums are the sums of the per-image partials, permuting the images leaves the loss
    unchanged, and a half batch reproduces its own partial sums."""
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512
    pri = PR.ssd512_canonical_priors()
    locs, scores, bx, lb = synth.make_train_batch(pri, 32, 81, 100, 1234 + 2)
    crit = MultiBoxLoss512(cu(pri, dev), cfg("", "", 81))
    l_d, s_d = cu(locs, dev), cu(scores, dev)
    bxd, lbd = [cu(b, dev) for b in bx], [cu(l, dev) for l in lb]
    loss = crit(l_d, s_d, bxd, lbd)
    st = crit.last["state"]
    part, sums = st.partials.cpu(), st.sums.cpu()
    assert torch.allclose(part.sum(0), sums, rtol=1e-12, atol=0)
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(3)).tolist()
    loss_p = crit(l_d[perm].contiguous(), s_d[perm].contiguous(), [bxd[i] for i in perm], [lbd[i] for i in perm])
    assert abs(loss_p.item() - loss.item()) <= 1e-6 * abs(loss.item())
    assert torch.equal(crit.last["state"].partials.cpu(), part[perm])
    crit(l_d[:16].contiguous(), s_d[:16].contiguous(), bxd[:16], lbd[:16])
    assert torch.equal(crit.last["state"].partials.cpu(), part[:16])
    # every image contributes at least one positive per object with overlap > 0 (forced match)
    assert float(part[:, 3].min()) >= 1


# ---------------------------------------------------------------------------------------------
# eval path
# ---------------------------------------------------------------------------------------------
def _run_detect(case, pri, locs, scores, keep, bt, dev):
    from shape_based_object_detection_b200.detect_scripts import detect_tools as DT
    from shape_based_object_detection_b200.models import utils as MU
    l_d, s_d, p_d = cu(locs, dev).contiguous(), cu(scores, dev), cu(pri, dev)
    k_d = cu(keep, dev) if keep is not None else None
    if case["fn"] == "utils.detect":
        out = MU.detect(l_d, s_d, case["min_score"], case["max_overlap"], case["top_k"], p_d,
                        cfg(n_classes=case["C"], box_type=bt, focal_type=case.get("focal_type", "softmax")),
                        prior_positives_idx=k_d)
    elif case["fn"] == "tools.detect":
        out = DT.detect(l_d, s_d, case["min_score"], case["max_overlap"], case["top_k"], p_d)
    else:
        out = DT.detect_refine(l_d, s_d, case["min_score"], case["max_overlap"], case["top_k"], p_d,
                               prior_positives_idx=k_d)
    return out, l_d


@pytest.mark.parametrize("name", list(DETECT_CASES))
def test_detect(golden, dev, name):
    case, G = DETECT_CASES[name], golden["detect"]
    pri, locs, scores, keep, bt = detect_inputs(case)
    (b, l, s), l_d = _run_detect(case, pri, locs, scores, keep, bt, dev)
    for i in range(case["N"]):
        want_l, want_s, want_b = T(G[f"{name}_l{i}"]).long(), T(G[f"{name}_s{i}"]), T(G[f"{name}_b{i}"])
        assert l[i].shape == want_l.shape, (l[i].shape, want_l.shape)
        assert torch.equal(l[i].cpu(), want_l)
        assert torch.allclose(s[i].cpu(), want_s, rtol=1e-5, atol=1e-8)
        assert torch.allclose(b[i].cpu(), want_b, rtol=1e-5, atol=1e-6)
    if bt == "corner":  # the reference clamps the caller's tensor in place
        assert float(l_d.min()) >= 0.0 and float(l_d.max()) <= 1.0


def test_detect_at_ssd512_coco_shape(dev):
    """P = 24 564, C = 81 (BASELINE config 2 eval), two images, against the oracle."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import utils as MU
    import torchvision
    pri = PR.ssd512_canonical_priors()
    locs, scores = synth.make_eval_batch(pri, 2, 81, 1234 + 2)
    want = O.detect(locs.clone(), scores, 0.01, 0.45, 200, pri, nms_fn=torchvision.ops.nms)
    got = MU.detect(cu(locs, dev), cu(scores, dev), 0.01, 0.45, 200, cu(pri, dev), cfg(n_classes=81))
    for i in range(2):
        assert torch.equal(got[1][i].cpu(), want[1][i])
        assert torch.allclose(got[2][i].cpu(), want[2][i], rtol=1e-5, atol=1e-8)
        assert torch.allclose(got[0][i].cpu(), want[0][i], rtol=1e-5, atol=1e-6)


def test_detect_full_batch_properties(dev):
    """N = 32 at the BASELINE shape: outputs are score-sorted, within [0,1], at most top_k, and the
    batch result equals the per-image results (images are independent)."""
    import shape_based_object_detection_b200 as S
    from shape_based_object_detection_b200 import priors as PR, synth
    pri = PR.ssd512_canonical_priors()
    locs, scores = synth.make_eval_batch(pri, 32, 81, 77)
    l_d, s_d, p_d = cu(locs, dev), cu(scores, dev), cu(pri, dev)
    ob, ol, osc, op, oc = S.detect_batched(l_d, s_d, 0.01, 0.45, 200, p_d)
    counts = oc.cpu()
    assert int(counts.min()) >= 1 and int(counts.max()) <= 200
    for i in (0, 7, 31):
        c = int(counts[i])
        sc = osc[i, :c].cpu()
        assert bool((sc[:-1] >= sc[1:]).all())
        assert float(ob[i, :c].min()) >= 0 and float(ob[i, :c].max()) <= 1
        b1, l1, s1, p1, c1 = S.detect_batched(l_d[i:i + 1].contiguous(), s_d[i:i + 1].contiguous(), 0.01, 0.45,
                                              200, p_d)
        assert int(c1[0]) == c
        assert torch.equal(l1[0, :c], ol[i, :c]) and torch.equal(p1[0, :c], op[i, :c])
        assert torch.equal(s1[0, :c], osc[i, :c])


def test_detect_speculative_cutoff_fallback(dev):
    """Every prior decodes to the same box, so NMS keeps one box per class and the top-k is never
    filled: the sampled score cutoff is too strict by construction and the exact fallback pass must
    reproduce the reference (class-major, all survivors)."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import utils as MU
    pri = PR.ssd300_priors()
    _, scores = synth.make_eval_batch(pri, 3, 7, 99, bg_bias=3.0)  # odd C: the fast (speculating) kernels
    locs = torch.tensor([0.2, 0.25, 0.6, 0.7]).repeat(3, pri.size(0), 1).contiguous()
    locs[2] += torch.rand((pri.size(0), 4), generator=torch.Generator().manual_seed(1)) * 0.3  # a normal image
    want = O.detect(locs.clone(), scores, 0.01, 0.45, 200, pri, box_type="corner")
    got = MU.detect(cu(locs.clone(), dev), cu(scores, dev), 0.01, 0.45, 200, cu(pri, dev),
                    cfg(n_classes=7, box_type="corner"))
    for i in range(3):
        assert torch.equal(got[1][i].cpu(), want[1][i]), i
        assert torch.allclose(got[2][i].cpu(), want[2][i], rtol=1e-5, atol=1e-8)
        assert torch.allclose(got[0][i].cpu(), want[0][i], rtol=1e-5, atol=1e-6)
    assert got[1][0].numel() <= 6
    # the workspace is clean again: a second call gives the same answer
    again = MU.detect(cu(locs.clone(), dev), cu(scores, dev), 0.01, 0.45, 200, cu(pri, dev),
                      cfg(n_classes=7, box_type="corner"))
    for i in range(3):
        assert torch.equal(again[1][i], got[1][i]) and torch.equal(again[2][i], got[2][i])


def test_detect_fast_kernels_all_variants(dev):
    """Odd class count (C = 21, the fast two-threads-per-row kernels with the speculative cutoff):
    per-class detect, top_k truncation, sigmoid scores, prior filter and the detect_tools second NMS,
    each against the oracle."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.detect_scripts import detect_tools as DT
    from shape_based_object_detection_b200.models import utils as MU
    pri = PR.ssd300_priors()
    locs, scores = synth.make_eval_batch(pri, 2, 21, 321, bg_bias=5.0)
    keep = scores[:, :, 1] > -1.0
    p_d = cu(pri, dev)

    def check(got, want):
        for i in range(2):
            assert torch.equal(got[1][i].cpu(), want[1][i])
            assert torch.allclose(got[2][i].cpu(), want[2][i], rtol=1e-5, atol=1e-8)
            assert torch.allclose(got[0][i].cpu(), want[0][i], rtol=1e-5, atol=1e-6)

    for top_k in (200, 17):
        check(MU.detect(cu(locs, dev), cu(scores, dev), 0.01, 0.45, top_k, p_d, cfg(n_classes=21)),
              O.detect(locs.clone(), scores, 0.01, 0.45, top_k, pri))
    check(MU.detect(cu(locs, dev), cu(scores, dev), 0.9, 0.45, 200, p_d, cfg(n_classes=21, focal_type="sigmoid")),
          O.detect(locs.clone(), scores, 0.9, 0.45, 200, pri, focal_type="sigmoid"))
    check(MU.detect(cu(locs, dev), cu(scores, dev), 0.01, 0.45, 200, p_d, cfg(n_classes=21),
                    prior_positives_idx=cu(keep, dev)),
          O.detect(locs.clone(), scores, 0.01, 0.45, 200, pri, prior_keep=keep))
    check(DT.detect(cu(locs, dev), cu(scores, dev), 0.01, 0.45, 200, p_d),
          O.detect(locs.clone(), scores, 0.01, 0.45, 200, pri, second_nms=0.7))


def test_forward_packed_and_host_entry(dev):
    """forward_packed (GT packed once) equals forward; the HOST-buffer C entry point
    sbod_loss_forward_host gives the same scalar."""
    import ctypes as C
    import shape_based_object_detection_b200 as S
    from shape_based_object_detection_b200 import _lib as L
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512
    pri = PR.ssd512_priors()
    locs, scores, bx, lb = synth.make_train_batch(pri, 4, 21, 12, 55)
    crit = MultiBoxLoss512(cu(pri, dev), cfg("", "", 21))
    bxd, lbd = [cu(b, dev) for b in bx], [cu(l, dev) for l in lb]
    l_d = cu(locs, dev).requires_grad_(True)
    s_d = cu(scores, dev).requires_grad_(True)
    a = crit(l_d, s_d, bxd, lbd)
    a.backward()
    ga, gb = l_d.grad.clone(), s_d.grad.clone()
    l_d.grad = s_d.grad = None
    packed = S.pack_ground_truth(bxd, lbd, dev)
    b = crit.forward_packed(l_d, s_d, packed)
    b.backward()
    assert a.item() == b.item()
    assert torch.equal(ga, l_d.grad) and torch.equal(gb, s_d.grad)
    # host entry point
    st = crit.last["state"]
    gt_b, gt_l, gt_o, gmax = [t.cpu() if hasattr(t, "cpu") else t for t in packed]
    h = L.LossDesc()
    keep_alive = [locs.contiguous(), scores.contiguous(), pri.contiguous(), crit.priors_xy.cpu().contiguous(),
                  gt_b.contiguous(), gt_l.contiguous(), gt_o.contiguous()]
    for name, t in zip(("locs", "scores", "priors_cxcy", "priors_xy", "gt_boxes", "gt_labels", "gt_offsets"), keep_alive):
        setattr(h, name, t.data_ptr())
    for f in ("N", "P", "C", "gmax", "thr_pos", "thr_neg", "reg_kind", "cls_kind", "binarize_labels", "neg_pos_ratio",
              "reg_weight", "smooth_l1_beta", "focal_alpha", "focal_gamma"):
        setattr(h, f, getattr(st.desc, f))
    T = int(gt_b.size(0))
    nbytes = L.lib().sbod_loss_forward_host_arena_bytes(C.byref(h), T)
    arena = torch.zeros(int(nbytes) + 256, dtype=torch.uint8, device=dev)
    out = torch.zeros(4, dtype=torch.float32).pin_memory()
    L.check(L.lib().sbod_loss_forward_host(C.byref(h), T, C.c_void_p(out.data_ptr()), L.ptr(arena),
                                           C.c_size_t(nbytes), L.stream_ptr()))
    torch.cuda.synchronize()
    assert abs(out[0].item() - a.item()) <= 1e-6 * abs(a.item())


def test_fcos_loss_and_postprocess(dev):
    """FCOS (parity unpinned w.r.t. the reference, pinned to the oracle's corrected restatement):
    label / box targets bit-exact, loss 1e-5, gradients 1e-4, post-process + detect vs the oracle."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import synth
    from shape_based_object_detection_b200.core import detect_batched, unpad_detections
    from shape_based_object_detection_b200.models import FCOSLoss, compute_location, fcos_postprocess
    locations = O.fcos_locations()
    mine = compute_location()
    assert all(torch.equal(a, b) for a, b in zip(locations, mine))
    P = sum(l.size(0) for l in locations)
    gen = torch.Generator().manual_seed(404)
    N, Cn = 3, 9
    bx, lb = synth.make_gt(N, 12, Cn, gen)
    bx[1] = torch.cat([bx[1], torch.tensor([[0.1, 0.1, 0.9, 0.9], [0.45, 0.45, 0.47, 0.47]])])  # large + tiny
    lb[1] = torch.cat([lb[1], torch.tensor([3, 5])])
    locs = torch.rand((N, P, 4), generator=gen) * 0.3 + 0.01
    scores = torch.randn((N, P, Cn), generator=gen)
    ctr = torch.randn((N, P), generator=gen)
    for cs in (True, False):
        l_c, s_c, c_c = [t.clone().requires_grad_(True) for t in (locs, scores, ctr)]
        want, parts = O.fcos_loss(locations, l_c, s_c, c_c, bx, lb, alpha=1.5, center_sample=cs, want_parts=True)
        want.backward()
        crit = FCOSLoss([cu(l, dev) for l in locations], cfg(n_classes=Cn) | {"reg_weights": 1.5},
                        center_sample=cs)
        l_d, s_d, c_d = [cu(t, dev).requires_grad_(True) for t in (locs, scores, ctr)]
        loss = crit(l_d, s_d, c_d, [cu(b, dev) for b in bx], [cu(l, dev) for l in lb])
        loss.backward()
        assert torch.equal(crit.last["labels"].cpu().long(), parts["labels"])
        posm = parts["labels"] > 0
        assert int(posm.sum()) > 20
        assert torch.equal(crit.last["targets"].cpu()[posm], parts["targets"][posm])
        assert abs(loss.item() - want.item()) <= RTOL_LOSS * abs(want.item()), (loss.item(), want.item())
        assert torch.allclose(l_d.grad.cpu(), l_c.grad, rtol=RTOL_GRAD, atol=1e-6)
        assert torch.allclose(s_d.grad.cpu(), s_c.grad, rtol=RTOL_GRAD, atol=1e-7)
        assert torch.allclose(c_d.grad.cpu(), c_c.grad, rtol=RTOL_GRAD, atol=1e-7)
    # eval side
    escores = scores * 2.0 - 2.0
    want_l, want_s = O.fcos_postprocess(locs, escores, ctr, locations)
    got_l, got_s = fcos_postprocess(cu(locs, dev), cu(escores, dev), cu(ctr, dev), [cu(l, dev) for l in locations])
    assert torch.allclose(got_l.cpu(), want_l, rtol=1e-6, atol=1e-7)
    assert torch.allclose(got_s.cpu(), want_s, rtol=1e-5, atol=1e-8)
    # detection on the oracle's probabilities (identical inputs on both sides), activation 'none'
    ref = O.detect(want_l.clone(), want_s, 0.05, 0.45, 100, None, box_type="corner", focal_type="none_is_identity")
    out = detect_batched(cu(want_l, dev).contiguous(), cu(want_s, dev), 0.05, 0.45, 100, None, act="none",
                         box_type="corner", clamp_inplace=True)
    got = unpad_detections(out[0], out[1], out[2], out[4])
    for i in range(N):
        assert torch.equal(got[1][i].cpu(), ref[1][i])
        assert torch.equal(got[2][i].cpu(), ref[2][i])
        assert torch.allclose(got[0][i].cpu(), ref[0][i], rtol=1e-6, atol=1e-7)


def test_ragged_shapes_and_many_objects(dev):
    """Edge cases of the streaming kernels: P not a multiple of the tile (or of 4: every tile start is
    mis-aligned, the last tile ends off a 16-byte boundary), a single image, an image with 300 objects
    (more than the per-warp object table), an image without objects next to normal ones."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512, RetinaFocalLoss
    from shape_based_object_detection_b200.models import utils as MU
    pri = PR.ssd512_priors()[:1003].contiguous()  # 1003 * 5 floats per image: odd everything
    gen = torch.Generator().manual_seed(9)
    for N, Cn, gmax, Mod, variant, reg, cls in ((3, 5, 8, MultiBoxLoss512, "s512", "", ""),
                                                (1, 7, 8, RetinaFocalLoss, "ret", "DIOU", "FOCAL"),
                                                (2, 5, 300, MultiBoxLoss512, "s512", "", "")):
        locs, scores, bx, lb = synth.make_train_batch(pri, N, Cn, gmax, 100 + N)
        if gmax == 300:
            bx[0], lb[0] = synth.make_gt(1, 300, Cn, gen, gmin=300)
            bx[0], lb[0] = bx[0][0], lb[0][0]
        l_c, s_c = locs.clone().requires_grad_(True), scores.clone().requires_grad_(True)
        want, parts = O.multibox_loss(variant, pri, l_c, s_c, bx, lb, reg_loss=reg, cls_loss=cls, want_parts=True)
        want.backward()
        crit = Mod(cu(pri, dev), cfg(reg, cls, Cn))
        l_d, s_d = cu(locs, dev).requires_grad_(True), cu(scores, dev).requires_grad_(True)
        loss = crit(l_d, s_d, [cu(b, dev) for b in bx], [cu(l, dev) for l in lb])
        loss.backward()
        st = crit.last["state"]
        assert torch.equal(st.obj.cpu().long(), parts["obj"]), (N, Cn, gmax)
        assert torch.equal(st.targets()[0].cpu(), parts["true_classes"])
        assert abs(loss.item() - want.item()) <= RTOL_LOSS * abs(want.item()), (loss.item(), want.item())
        assert torch.allclose(l_d.grad.cpu(), l_c.grad, rtol=RTOL_GRAD, atol=1e-7)
        assert torch.allclose(s_d.grad.cpu(), s_c.grad, rtol=RTOL_GRAD, atol=1e-7)
    # an image without objects: defined as "all background" (the reference raises); the other images
    # must be unaffected, i.e. the batch sums equal those of the batch without that image
    locs, scores, bx, lb = synth.make_train_batch(pri, 3, 5, 8, 321)
    crit = MultiBoxLoss512(cu(pri, dev), cfg("", "", 5))
    bx2 = [bx[0], torch.zeros((0, 4)), bx[2]]
    lb2 = [lb[0], torch.zeros((0,), dtype=torch.long), lb[2]]
    crit(cu(locs, dev), cu(scores, dev), [cu(b, dev) for b in bx2], [cu(l, dev) for l in lb2])
    part = crit.last["state"].partials.cpu()
    assert float(part[1, 3]) == 0.0 and float(part[1, 0]) == 0.0 and float(part[1, 2]) == 0.0
    crit(cu(locs[[0, 2]], dev).contiguous(), cu(scores[[0, 2]], dev).contiguous(),
         [cu(bx[0], dev), cu(bx[2], dev)], [cu(lb[0], dev), cu(lb[2], dev)])
    assert torch.equal(crit.last["state"].partials.cpu(), part[[0, 2]])
    # eval path on the same ragged shape
    elocs, escores = synth.make_eval_batch(pri, 3, 5, 11, bg_bias=3.0)
    want = O.detect(elocs.clone(), escores, 0.01, 0.45, 50, pri)
    got = MU.detect(cu(elocs, dev), cu(escores, dev), 0.01, 0.45, 50, cu(pri, dev), cfg(n_classes=5))
    for i in range(3):
        assert torch.equal(got[1][i].cpu(), want[1][i])
        assert torch.allclose(got[2][i].cpu(), want[2][i], rtol=1e-5, atol=1e-8)


def test_fast_division_is_correctly_rounded(dev):
    """The matching kernel divides with its own reciprocal + FMA sequence (csrc/common.cuh:div_rn_fast)
    wherever div_fast_ok holds; it must equal IEEE division bit for bit there."""
    import ctypes as C
    from shape_based_object_detection_b200 import _lib as L
    gen = torch.Generator().manual_seed(77)
    n = 1 << 22
    # IoU-shaped operands, wide-exponent operands, zeros and tiny numerators
    inter = torch.rand(n, generator=gen) * torch.rand(n, generator=gen)
    union = inter + torch.rand(n, generator=gen) + 1e-5
    a = torch.cat([inter, torch.exp2(torch.rand(n, generator=gen) * 130 - 65), torch.zeros(1024),
                   torch.rand(1024, generator=gen) * 1e-30])
    b = torch.cat([union, torch.exp2(torch.rand(n, generator=gen) * 130 - 65), torch.rand(1024, generator=gen) + 0.1,
                   torch.rand(1024, generator=gen) + 0.1])
    a_d, b_d = a.to(dev), b.to(dev)
    out, ref = torch.empty_like(a_d), torch.empty_like(a_d)
    ok = torch.empty(a_d.numel(), dtype=torch.uint8, device=dev)
    L.check(L.lib().sbod_selftest_div(L.ptr(a_d), L.ptr(b_d), C.c_longlong(a_d.numel()), L.ptr(out), L.ptr(ref),
                                      L.ptr(ok), L.stream_ptr()))
    torch.cuda.synchronize()
    okb = ok.bool()
    assert okb[:n].all() and okb[2 * n:2 * n + 1024].all()      # IoU-shaped operands and zeros take the fast sequence
    assert not okb[2 * n + 1024:].any()                         # tiny numerators do not
    assert torch.equal(out[okb].view(torch.int32), ref[okb].view(torch.int32))
    assert torch.equal(ref.cpu(), a / b)                        # div.rn itself == torch's division


@pytest.mark.parametrize("shape", ["ssd512_32", "retina640_6", "many_objects", "one_object"])
def test_fused_assignment_equals_generic_kernel_at_full_size(dev, shape):
    """The warp-specialised match role (ticket queues, staged object pairs, register keys) against the
    generic one-thread-per-prior assignment kernel (sbod_assign, itself pinned to the oracle on small
    cases) at sizes the CPU oracle cannot do: overlaps and object indices after the forced-match
    override must be bit-identical, whatever the ticket geometry."""
    import shape_based_object_detection_b200 as S
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.dataset.transforms import cxcy_to_xy
    from shape_based_object_detection_b200.models import MultiBoxLoss512
    name, N, gmax, gmin = {"ssd512_32": ("ssd512_canonical", 32, 100, 1), "retina640_6": ("retinanet640", 6, 100, 1),
                           "many_objects": ("ssd512_canonical", 3, 300, 129),
                           "one_object": ("ssd512_canonical", 8, 1, 1)}[shape]
    pri = PR.PRIOR_TABLES[name]()
    gen = torch.Generator().manual_seed({"ssd512_32": 41, "retina640_6": 42, "many_objects": 43, "one_object": 44}[shape])
    bx, lb = synth.make_gt(N, gmax, 81, gen, gmin=gmin)
    P = pri.size(0)
    locs = torch.randn((N, P, 4), generator=gen) * 0.1
    scores = torch.randn((N, P, 81), generator=gen)
    pri_d = cu(pri, dev)
    bxd, lbd = [cu(b, dev) for b in bx], [cu(l, dev) for l in lb]
    crit = MultiBoxLoss512(pri_d, cfg("", "", 81))
    crit(cu(locs, dev), cu(scores, dev), bxd, lbd)
    st = crit.last["state"]
    ov, obj, cls, neg = S.assign(bxd, lbd, cxcy_to_xy(pri_d), threshold=0.5)
    assert torch.equal(st.obj, obj)
    assert torch.equal(st.ov.view(torch.int32), ov.view(torch.int32))
    cls_f, neg_f = st.targets()
    assert torch.equal(cls_f, cls)


def test_split_phase_detect_equals_detect_batched(dev):
    """detect_begin (sampling pass on a side stream) + detect_end == detect_batched."""
    import shape_based_object_detection_b200 as S
    from shape_based_object_detection_b200 import priors as PR, synth
    pri = PR.ssd512_canonical_priors()
    elocs, escores = synth.make_eval_batch(pri, 4, 81, 99)
    l_d, s_d, p_d = cu(elocs, dev), cu(escores, dev), cu(pri, dev)
    want = S.detect_batched(l_d, s_d, 0.01, 0.45, 200, p_d)
    side = torch.cuda.Stream(device=dev)
    call = S.detect_begin(l_d, s_d, 0.01, 0.45, 200, p_d, side_stream=side)
    torch.cuda._sleep(200000)  # unrelated work on the main stream in between
    got = S.detect_end(call)
    torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------
# metrics.calculate_mAP: the consumer of detect()'s output
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["small", "voc_like", "strict", "sparse"])
def test_map_matches_reference_fixture(golden, dev, name):
    from cases import MAP_CASES, map_inputs
    from shape_based_object_detection_b200.metrics import calculate_mAP
    case = MAP_CASES[name]
    label_map = {("background" if i == 0 else "c%d" % i): i for i in range(case["n_classes"])}
    args = [[cu(t, dev) for t in lst] for lst in map_inputs(case)]
    aps, mean_ap = calculate_mAP(*args, case["threshold"], label_map, device="cuda:0")
    got = torch.tensor([aps["c%d" % i] for i in range(1, case["n_classes"])], dtype=torch.float32)
    want = T(golden["map"][name + "_ap"])
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), (got - want).abs().max()
    assert abs(mean_ap - float(golden["map"][name + "_map"])) <= 1e-6


def test_map_large_and_edge_cases(dev):
    """Against the oracle at a size the Python reference needs minutes for, plus: no detections at all,
    images without objects, a class that only has difficult objects."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import synth
    from shape_based_object_detection_b200.metrics import calculate_mAP
    n_classes = 21
    label_map = {("background" if i == 0 else "c%d" % i): i for i in range(n_classes)}
    case = list(synth.make_map_case(150, n_classes, 15, 200, 77, ties=True))
    case[3][5] = torch.zeros((0, 4))                      # an image without objects (its detections are FPs)
    case[4][5] = torch.zeros((0,), dtype=torch.int64)
    case[5][5] = torch.zeros((0,), dtype=torch.uint8)
    for i in range(len(case[4])):                         # class 3 only has difficult objects
        case[5][i] = torch.where(case[4][i] == 3, torch.ones_like(case[5][i]), case[5][i])
    want_ap, want_map = O.calculate_mAP(*case, 0.5, n_classes)
    aps, mean_ap = calculate_mAP(*[[cu(t, dev) for t in lst] for lst in case], 0.5, label_map)
    got = torch.tensor([aps["c%d" % i] for i in range(1, n_classes)], dtype=torch.float32)
    assert torch.allclose(got, want_ap, rtol=1e-6, atol=1e-7), (got - want_ap).abs().max()
    assert abs(mean_ap - want_map) <= 1e-6
    assert aps["c3"] == 0.0 and aps["c20"] == 0.0
    # no detections
    empty = [[torch.zeros((0, 4), device=dev)] * 3, [torch.zeros((0,), dtype=torch.int64, device=dev)] * 3,
             [torch.zeros((0,), device=dev)] * 3]
    gt = [[cu(t, dev) for t in lst[:3]] for lst in case[3:]]
    aps, mean_ap = calculate_mAP(*empty, *gt, 0.5, label_map)
    assert mean_ap == 0.0 and all(v == 0.0 for v in aps.values())


def test_bce_focal_loss(golden, dev):
    """operators.Loss.FocalLoss (Loss.py:83-103) against the reference fixture, value and gradient."""
    from shape_based_object_detection_b200.operators import Loss as LS
    I, G = operator_inputs(), golden["extras"]
    for tag, scale in (("", 1.0), ("_wide", 6.0)):
        x = cu(I["lg"] * scale, dev).requires_grad_(True)
        fl = LS.FocalLoss(0.25, 2)(x, cu(I["tg"], dev))
        fl.backward()
        want = float(G["bcefocal" + tag])
        assert abs(fl.item() - want) <= RTOL_LOSS * abs(want), (fl.item(), want)
        assert torch.allclose(x.grad.cpu(), T(G["bcefocal" + tag + "_g"]), rtol=RTOL_GRAD, atol=1e-6)


def test_diounms(golden, dev):
    """operators.iou_utils.diounms against the reference fixture, and a bigger random problem against
    the oracle."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200.operators import iou_utils as U
    I, G = operator_inputs(), golden["extras"]
    for beta in (1.0, 0.6):
        k, cnt = U.diounms(cu(I["nb"], dev), cu(I["ns"], dev), 0.45, 200, beta)
        assert torch.equal(k[:cnt].cpu(), T(G["diounms_b%02d" % int(beta * 10)]).long())
        assert int(k[cnt:].abs().sum()) == 0
    g = torch.Generator().manual_seed(15)
    c = torch.rand((2000, 2), generator=g)
    wh = torch.rand((2000, 2), generator=g) * 0.2 + 0.01
    b = torch.cat([c - wh / 2, c + wh / 2], 1)
    s = torch.rand((2000,), generator=g)
    want, wcnt = O.diounms(b, s, 0.4, 500, 1.0)
    k, cnt = U.diounms(cu(b, dev), cu(s, dev), 0.4, 500, 1.0)
    assert cnt == wcnt and torch.equal(k[:cnt].cpu(), want[:wcnt])
