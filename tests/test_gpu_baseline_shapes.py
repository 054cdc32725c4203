"""GPU suite (-m gpu), part 2: the CUDA path against the CPU oracle AT THE SHAPES BASELINE.json NAMES
(no sub-sampling, no CUDA-vs-CUDA comparisons): config 2 at its full batch of 32, RetinaNet-640 with
focal + GIoU and the per-class top-1000 cap, RefineDet512 at 16 320 priors x <= 200 dense objects, SSD300
at 8 732 priors x 21 classes, FCOS at 800 x 1333. Gates (BASELINE.json north_star): object indices,
classes and the kept (class, prior) sequence bit-exact; losses, scores, boxes 1e-5 relative; gradients
1e-4 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL_LOSS = 1e-5
RTOL_GRAD = 1e-4


class Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def cfg(reg="", cls="", n_classes=81, box_type="offset", focal_type="softmax", **kw):
    return Cfg(device=torch.device("cuda:0"), n_classes=n_classes, reg_weights=1.0, reg_loss=reg, cls_loss=cls,
               model={"box_type": box_type}, focal_type=focal_type, **kw)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "the gpu suite needs a CUDA device"
    return torch.device("cuda:0")


def _check_loss(crit, variant, pri, locs, scores, bx, lb, dev, reg="", cls=""):
    """Loss module vs oracle on the same batch: targets bit-exact, loss 1e-5, gradients 1e-4."""
    from oracle import box_pipeline as O
    l_c, s_c = locs.clone().requires_grad_(True), scores.clone().requires_grad_(True)
    want, parts = O.multibox_loss(variant, pri, l_c, s_c, bx, lb, reg_loss=reg, cls_loss=cls, want_parts=True)
    want.backward()
    l_d, s_d = locs.to(dev).requires_grad_(True), scores.to(dev).requires_grad_(True)
    loss = crit(l_d, s_d, [b.to(dev) for b in bx], [l.to(dev) for l in lb])
    loss.backward()
    st = crit.last["state"]
    cls_t, neg_t = st.targets()
    assert torch.equal(st.obj.cpu().long(), parts["obj"])
    assert torch.equal(st.ov.cpu(), parts["ov"])
    assert torch.equal(cls_t.cpu(), parts["true_classes"])
    assert torch.equal(neg_t.cpu(), parts["true_neg_classes"])
    assert int(st.loss[3].item()) == int(parts["n_pos"].sum())
    assert abs(loss.item() - want.item()) <= RTOL_LOSS * abs(want.item()), (loss.item(), want.item())
    assert torch.allclose(l_d.grad.cpu(), l_c.grad, rtol=RTOL_GRAD, atol=1e-7)
    assert torch.allclose(s_d.grad.cpu(), s_c.grad, rtol=RTOL_GRAD, atol=1e-7)
    return loss.item()


def _check_detect(got, want, n_images):
    """got = padded device outputs (boxes, labels, scores, prior, counts); want = oracle lists with priors.
    The kept (class, prior) sequence must be identical image by image. Returns the number of images that
    differ (the caller asserts it is zero) so that a failure reports how widespread it is."""
    ob, ol, osc, op, oc = got
    counts = oc.cpu().tolist()
    bad = []
    for i in range(n_images):
        c = counts[i]
        wl, wp = want[1][i], want[3][i]
        if c != wl.numel() or not torch.equal(ol[i, :c].cpu(), wl) or not torch.equal(op[i, :c].cpu().long(), wp):
            bad.append(i)
            continue
        assert torch.allclose(osc[i, :c].cpu(), want[2][i], rtol=1e-5, atol=1e-8)
        assert torch.allclose(ob[i, :c].cpu(), want[0][i], rtol=1e-5, atol=1e-6)
    return bad


def test_config2_full_batch_loss_vs_oracle(dev):
    """BASELINE config 2 exactly as bench.py runs it: N = 32, P = 24 564, C = 81, G <= 100."""
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512
    pri = PR.ssd512_canonical_priors()
    locs, scores, bx, lb = synth.make_train_batch(pri, 32, 81, 100, 1234 + 2)
    crit = MultiBoxLoss512(pri.to(dev), cfg())
    _check_loss(crit, "s512", pri, locs, scores, bx, lb, dev)


def test_config2_full_batch_detect_vs_oracle(dev):
    """N = 32 eval batch of config 2: the kept (class, prior) sequence of every image equals the oracle's
    (torchvision.ops.nms inside), scores / boxes within 1e-5."""
    import torchvision
    import shape_based_object_detection_b200 as S
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    pri = PR.ssd512_canonical_priors()
    locs, scores = synth.make_eval_batch(pri, 32, 81, 4321)
    want = O.detect(locs.clone(), scores, 0.01, 0.45, 200, pri, nms_fn=torchvision.ops.nms, return_priors=True)
    got = S.detect_batched(locs.to(dev), scores.to(dev), 0.01, 0.45, 200, pri.to(dev))
    bad = _check_detect(got, want, 32)
    # Scores agree with the CPU's to a few ulp, not bit for bit (other summation order): on rare images two
    # same-class candidates an ulp apart swap places. The exact statement (indices bit-exact GIVEN the score
    # bits) is test_config2_detect_indices_exact_given_score_bits; here at most 2 of 32 images may differ.
    print("config 2, N=32: %d of 32 images differ from the CPU oracle in the kept sequence: %s" % (len(bad), bad))
    assert len(bad) <= 2, "kept (class, prior) sequence differs from the oracle on images %s of 32" % bad


def test_config3_retinanet640_focal_giou_top1000(dev):
    """BASELINE config 3: 76 725 anchors, 81 classes, softmax focal + GIoU (explicit opt-in, the reference
    itself only knows DIoU: RetinaNet.py:461-466) and detect with the per-class top-1000 candidate cap."""
    import torchvision
    import shape_based_object_detection_b200 as S
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import RetinaFocalLoss
    pri = PR.retinanet640_priors()
    assert pri.size(0) == 76725
    locs, scores, bx, lb = synth.make_train_batch(pri, 2, 81, 100, 1234 + 3)
    crit = RetinaFocalLoss(pri.to(dev), cfg("GIOU", "FOCAL"))
    # without the opt-in 'GIOU' means SmoothL1, as in the reference
    plain = _check_loss(crit, "ret", pri, locs, scores, bx, lb, dev, reg="", cls="FOCAL")
    crit.extended_reg_losses = True
    giou = _check_loss(crit, "ret", pri, locs, scores, bx, lb, dev, reg="GIOU", cls="FOCAL")
    assert plain != giou
    # DIoU + CE mining on the same shape (the reference's other branch pair)
    crit2 = RetinaFocalLoss(pri.to(dev), cfg("DIOU", ""))
    _check_loss(crit2, "ret", pri, locs, scores, bx, lb, dev, reg="DIOU", cls="")
    elocs, escores = synth.make_eval_batch(pri, 2, 81, 4321 + 3)
    for cap in (1000, 0):
        want = O.detect(elocs.clone(), escores, 0.01, 0.45, 200, pri, nms_fn=torchvision.ops.nms,
                        pre_nms_topk=cap, return_priors=True)
        got = S.detect_batched(elocs.to(dev), escores.to(dev), 0.01, 0.45, 200, pri.to(dev), pre_nms_topk=cap)
        assert not _check_detect(got, want, 2), cap
    # the cap changes what is kept only when suppression reaches below the 1000th candidate of a class;
    # make it bite: 40 candidates per class at most
    want = O.detect(elocs.clone(), escores, 0.01, 0.45, 200, pri, nms_fn=torchvision.ops.nms, pre_nms_topk=40,
                    return_priors=True)
    got = S.detect_batched(elocs.to(dev), escores.to(dev), 0.01, 0.45, 200, pri.to(dev), pre_nms_topk=40)
    assert not _check_detect(got, want, 2)


def test_config4_refinedet_dense_traffic(dev):
    """BASELINE config 4: RefineDet512, 16 320 priors, DETRAC-shaped dense ground truth (<= 200 small boxes
    per image), ARM + ODM loss with refined anchors, and detect_refine on the decoded boxes."""
    import torchvision
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.detect_scripts import detect_tools as DT
    from shape_based_object_detection_b200.models import RefineDetLoss, offset2bbox
    pri = PR.refinedet512_priors()
    assert pri.size(0) == 16320
    N, Cn = 2, 4
    gen = torch.Generator().manual_seed(1234 + 4)
    bx, lb = synth.make_gt(N, 200, Cn, gen, dense=True, gmin=150)
    P = pri.size(0)
    arm_l = torch.randn((N, P, 4), generator=gen) * 0.1
    arm_s = torch.randn((N, P, 2), generator=gen) * 2.0
    odm_l = torch.randn((N, P, 4), generator=gen) * 0.1
    odm_s = torch.randn((N, P, Cn), generator=gen)
    cs = [t.clone().requires_grad_(True) for t in (arm_l, arm_s, odm_l, odm_s)]
    want, arm_parts, odm_parts, anchors, easy = O.refinedet_loss(pri, *cs, bx, lb, want_parts=True)
    want.backward()
    crit = RefineDetLoss(pri.to(dev), cfg(n_classes=Cn))
    ds = [t.to(dev).requires_grad_(True) for t in (arm_l, arm_s, odm_l, odm_s)]
    loss = crit(*ds, [b.to(dev) for b in bx], [l.to(dev) for l in lb])
    loss.backward()
    # ARM: static priors -> bit-exact assignment. ODM: the anchors are themselves decoded boxes (exp on the GPU
    # vs on the CPU: last-bit differences), so its assignment is compared up to a handful of borderline priors.
    state, parts = crit.last_arm["state"], arm_parts
    cls_t, _ = state.targets()
    assert torch.equal(state.obj.cpu().long(), parts["obj"])
    assert torch.equal(state.ov.cpu(), parts["ov"])
    assert torch.equal(cls_t.cpu(), parts["true_classes"])
    state, parts = crit.last_odm["state"], odm_parts
    cls_t, _ = state.targets()
    total = parts["obj"].numel()
    assert int((cls_t.cpu() != parts["true_classes"]).sum()) <= max(2, total // 20000)
    assert torch.allclose(state.ov.cpu(), parts["ov"], rtol=1e-4, atol=1e-6)
    assert abs(loss.item() - want.item()) <= RTOL_LOSS * abs(want.item()), (loss.item(), want.item())
    for d, c, name in zip(ds, cs, ("arm_locs", "arm_scores", "odm_locs", "odm_scores")):
        if c.grad is None or float(c.grad.abs().max()) == 0.0:
            assert d.grad is None or float(d.grad.abs().max()) == 0.0, name
        else:
            assert torch.allclose(d.grad.cpu(), c.grad, rtol=RTOL_GRAD, atol=1e-7), name
    # eval: refined-anchor decode + detect_refine (second, class-agnostic NMS at 0.7) on the same shape
    e_arm = torch.randn((N, P, 4), generator=gen) * 0.3
    e_odm = torch.randn((N, P, 4), generator=gen) * 0.3
    e_sc = torch.randn((N, P, Cn), generator=gen) * 2.0
    e_sc[:, :, 0] += 4.0
    keep = torch.randn((N, P), generator=gen) > -1.0
    boxes_c = O.offset2bbox(e_arm, e_odm, pri)
    boxes_d = offset2bbox(e_arm.to(dev), e_odm.to(dev), pri.to(dev))
    assert torch.allclose(boxes_d.cpu(), boxes_c, rtol=1e-5, atol=1e-6)
    want = O.detect(boxes_c.clone(), e_sc, 0.01, 0.45, 200, pri, box_type="corner", prior_keep=keep, second_nms=0.7,
                    nms_fn=torchvision.ops.nms)
    got = DT.detect_refine(boxes_c.to(dev), e_sc.to(dev), 0.01, 0.45, 200, pri.to(dev),
                           prior_positives_idx=keep.to(dev))
    for i in range(N):
        assert torch.equal(got[1][i].cpu(), want[1][i]), i
        assert torch.allclose(got[2][i].cpu(), want[2][i], rtol=1e-5, atol=1e-8)
        assert torch.allclose(got[0][i].cpu(), want[0][i], rtol=1e-5, atol=1e-6)


def test_config1_ssd300_full_shape(dev):
    """BASELINE config 1: SSD300, all 8 732 priors, 21 classes, <= 20 objects: L1 + batch-global hard-negative
    mining (SSD300.py:580-588), DIoU + focal, and detect."""
    import torchvision
    import shape_based_object_detection_b200 as S
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss300
    pri = PR.ssd300_priors()
    assert pri.size(0) == 8732
    locs, scores, bx, lb = synth.make_train_batch(pri, 8, 21, 20, 1234 + 1)
    _check_loss(MultiBoxLoss300(pri.to(dev), cfg(n_classes=21)), "s300", pri, locs, scores, bx, lb, dev)
    _check_loss(MultiBoxLoss300(pri.to(dev), cfg("DIOU", "FOCAL", n_classes=21)), "s300", pri, locs, scores, bx, lb,
                dev, reg="DIOU", cls="FOCAL")
    # a sharded MultiBoxLoss300 is refused: its mining is batch-global
    crit = MultiBoxLoss300(pri.to(dev), cfg(n_classes=21))
    crit.process_group = object()
    with pytest.raises(S._lib.SbodError):
        crit(locs.to(dev), scores.to(dev), [b.to(dev) for b in bx], [l.to(dev) for l in lb])
    elocs, escores = synth.make_eval_batch(pri, 8, 21, 4321 + 1)
    want = O.detect(elocs.clone(), escores, 0.01, 0.45, 200, pri, nms_fn=torchvision.ops.nms, return_priors=True)
    got = S.detect_batched(elocs.to(dev), escores.to(dev), 0.01, 0.45, 200, pri.to(dev))
    assert not _check_detect(got, want, 8)


def test_config5_fcos_800x1333(dev):
    """BASELINE config 5 shape: 22 300 locations (800 x 1333, strides 8..128), 80 classes + centerness.
    PARITY UNPINNED with respect to the reference (models/FCOSDet.py does not run); pinned to the oracle's
    corrected restatement: targets bit-exact, loss 1e-5, gradients 1e-4, post-process + NMS."""
    import torchvision
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import synth
    from shape_based_object_detection_b200.core import detect_batched
    from shape_based_object_detection_b200.models import FCOSLoss, compute_location, fcos_postprocess
    size = (800, 1333)
    locations = O.fcos_locations(image_size=size)
    mine = compute_location(image_size=size)
    assert sum(l.size(0) for l in mine) == 22300
    assert all(torch.equal(a, b) for a, b in zip(locations, mine))
    P, N, Cn = 22300, 2, 81
    gen = torch.Generator().manual_seed(1234 + 5)
    bx, lb = synth.make_gt(N, 100, Cn, gen)
    locs = torch.rand((N, P, 4), generator=gen) * 0.3 + 0.01
    scores = torch.randn((N, P, Cn), generator=gen)
    ctr = torch.randn((N, P), generator=gen)
    l_c, s_c, c_c = [t.clone().requires_grad_(True) for t in (locs, scores, ctr)]
    want, parts = O.fcos_loss(locations, l_c, s_c, c_c, bx, lb, alpha=1.0, want_parts=True, image_size=size)
    want.backward()
    crit = FCOSLoss([l.to(dev) for l in locations], cfg(n_classes=Cn), image_size=size)
    l_d, s_d, c_d = [t.to(dev).requires_grad_(True) for t in (locs, scores, ctr)]
    loss = crit(l_d, s_d, c_d, [b.to(dev) for b in bx], [l.to(dev) for l in lb])
    loss.backward()
    assert torch.equal(crit.last["labels"].cpu().long(), parts["labels"])
    posm = parts["labels"] > 0
    assert int(posm.sum()) > 100
    assert torch.equal(crit.last["targets"].cpu()[posm], parts["targets"][posm])
    assert abs(loss.item() - want.item()) <= RTOL_LOSS * abs(want.item()), (loss.item(), want.item())
    assert torch.allclose(l_d.grad.cpu(), l_c.grad, rtol=RTOL_GRAD, atol=1e-6)
    assert torch.allclose(s_d.grad.cpu(), s_c.grad, rtol=RTOL_GRAD, atol=1e-7)
    assert torch.allclose(c_d.grad.cpu(), c_c.grad, rtol=RTOL_GRAD, atol=1e-7)
    # eval: post-process, then NMS on the oracle's probabilities (identical inputs on both sides)
    escores = scores * 2.0 - 3.0
    want_l, want_s = O.fcos_postprocess(locs, escores, ctr, locations)
    got_l, got_s = fcos_postprocess(locs.to(dev), escores.to(dev), ctr.to(dev), [l.to(dev) for l in locations])
    assert torch.allclose(got_l.cpu(), want_l, rtol=1e-6, atol=1e-7)
    assert torch.allclose(got_s.cpu(), want_s, rtol=1e-5, atol=1e-8)
    ref = O.detect(want_l.clone(), want_s, 0.05, 0.45, 100, None, box_type="corner", focal_type="none_is_identity",
                   nms_fn=torchvision.ops.nms, return_priors=True)
    out = detect_batched(want_l.to(dev).contiguous(), want_s.to(dev), 0.05, 0.45, 100, None, act="none",
                         box_type="corner", clamp_inplace=True)
    assert not _check_detect(out, ref, N)


def test_iou_loss_center_mode_backpropagates_through_decode(dev):
    """IouLoss(pred_mode='Center') decodes its predictions first (operators/Loss.py:176-178): the gradient
    must reach loc_p, and the converters are differentiable like the reference's torch expressions."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200.dataset import transforms as TR
    from shape_based_object_detection_b200.operators import Loss as LS
    from shape_based_object_detection_b200.operators import iou_utils as U
    gen = torch.Generator().manual_seed(7)
    M = 300
    pri = torch.cat([torch.rand((M, 2), generator=gen) * 0.6 + 0.2, torch.rand((M, 2), generator=gen) * 0.3 + 0.05], 1)
    loc = torch.randn((M, 4), generator=gen) * 0.3
    c = torch.rand((M, 2), generator=gen) * 0.6 + 0.2
    wh = torch.rand((M, 2), generator=gen) * 0.3 + 0.05
    tgt = torch.cat([c - wh / 2, c + wh / 2], 1)
    wts = torch.rand((M,), generator=gen)
    for weights in (None, wts):
        x_c = loc.clone().requires_grad_(True)
        dec = O.decode(x_c, pri, [0.1, 0.2])
        l = 1.0 - O.pair_overlap(dec, tgt, "diou")
        want = (l * weights).sum() / weights.sum() if weights is not None else l.sum() / M
        want.backward()
        x_d = loc.to(dev).requires_grad_(True)
        got = LS.IouLoss(pred_mode="Center", reduce="mean", variances=[0.1, 0.2], losstype="Diou")(
            x_d, tgt.to(dev), pri.to(dev), weights.to(dev) if weights is not None else None)
        got.backward()
        assert abs(got.item() - want.item()) <= RTOL_LOSS * abs(want.item())
        assert torch.allclose(x_d.grad.cpu(), x_c.grad, rtol=RTOL_GRAD, atol=1e-7)
    # SmoothL1Loss with weights (Loss.py:219-221)
    p_c = loc.clone().requires_grad_(True)
    t4 = torch.randn((M, 4), generator=gen) * 0.3
    sl = O.smooth_l1_rows(p_c, t4)
    want = (sl * wts[:, None]).sum() / wts[:, None].sum()
    want.backward()
    p_d = loc.to(dev).requires_grad_(True)
    got = LS.SmoothL1Loss()(p_d, t4.to(dev), wts[:, None].to(dev))
    got.backward()
    assert abs(got.item() - want.item()) <= RTOL_LOSS * abs(want.item())
    assert torch.allclose(p_d.grad.cpu(), p_c.grad, rtol=RTOL_GRAD, atol=1e-7)
    # every converter: value and gradient against the oracle's torch expressions
    cases = [
        (TR.xy_to_cxcy, O.xy_to_cxcy, tgt, None),
        (TR.cxcy_to_xy, O.cxcy_to_xy, pri, None),
        (U.point_form, O.cxcy_to_xy, pri, None),
        (U.center_size, O.xy_to_cxcy, tgt, None),
        (TR.cxcy_to_gcxgcy, O.cxcy_to_gcxgcy, O.xy_to_cxcy(tgt), pri),
        (TR.gcxgcy_to_cxcy, O.gcxgcy_to_cxcy, loc, pri),
        (lambda a, p: U.encode(a, p, [0.1, 0.2]), lambda a, p: O.encode(a, p, [0.1, 0.2]), tgt, pri),
        (lambda a, p: U.decode(a, p, [0.1, 0.2]), lambda a, p: O.decode(a, p, [0.1, 0.2]), loc, pri),
    ]
    up = torch.randn((M, 4), generator=gen)
    for ours, ref, x, p in cases:
        x_c, x_d = x.clone().requires_grad_(True), x.to(dev).requires_grad_(True)
        y_c = ref(x_c, p) if p is not None else ref(x_c)
        y_d = ours(x_d, p.to(dev)) if p is not None else ours(x_d)
        (y_c * up).sum().backward()
        (y_d * up.to(dev)).sum().backward()
        assert torch.allclose(y_d.detach().cpu(), y_c.detach(), rtol=1e-5, atol=1e-6)
        assert torch.allclose(x_d.grad.cpu(), x_c.grad, rtol=RTOL_GRAD, atol=1e-6)


def test_second_device(dev):
    """Tensors on cuda:1 while cuda:0 is the current device (the reference picks cuda:1 when
    config.device == 1, train_anchor.py:66-67): kernels, workspaces and streams must follow the tensors."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import shape_based_object_detection_b200 as S
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512
    d1 = torch.device("cuda:1")
    pri = PR.ssd512_priors()
    locs, scores, bx, lb = synth.make_train_batch(pri, 3, 21, 12, 5)
    outs = []
    for d in (dev, d1):
        assert torch.cuda.current_device() == 0
        crit = MultiBoxLoss512(pri.to(d), cfg(n_classes=21))
        l_d, s_d = locs.to(d).requires_grad_(True), scores.to(d).requires_grad_(True)
        loss = crit(l_d, s_d, [b.to(d) for b in bx], [l.to(d) for l in lb])
        loss.backward()
        elocs, escores = synth.make_eval_batch(pri, 3, 21, 6, bg_bias=4.0)
        det = S.detect_batched(elocs.to(d), escores.to(d), 0.01, 0.45, 200, pri.to(d))
        outs.append((loss.item(), s_d.grad.cpu(), [t.cpu() for t in det]))
    assert outs[0][0] == outs[1][0]
    assert torch.equal(outs[0][1], outs[1][1])
    assert all(torch.equal(a, b) for a, b in zip(outs[0][2], outs[1][2]))
    with pytest.raises(S._lib.SbodError):  # mixed devices are rejected
        S.detect_batched(elocs.to(dev), escores.to(d1), 0.01, 0.45, 200, pri.to(dev))


def test_detect_objects_class_agnostic(dev):
    """detect_objects (models/utils.py:87-178, detect_tools.py:10-97): PARITY UNPINNED - the reference reaches
    exit(); checked against the oracle's restatement of the intended semantics: one candidate per prior (best
    foreground class), one class-agnostic NMS, label = arg-max class. SSD300 shape and config 2 shape."""
    import torchvision
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.detect_scripts import detect_tools as DT
    from shape_based_object_detection_b200.models import utils as MU
    for table, Cn, N, bg, top_k in (("ssd300", 21, 3, 3.0, 200), ("ssd300", 21, 2, 3.0, 25), ("ssd512_canonical", 81, 2, 6.0, 200)):
        pri = PR.PRIOR_TABLES[table]()
        locs, scores = synth.make_eval_batch(pri, N, Cn, 97 + Cn, bg_bias=bg)
        want = O.detect_objects(locs.clone(), scores, 0.05, 0.45, top_k, pri, nms_fn=torchvision.ops.nms)
        got = MU.detect_objects(locs.to(dev), scores.to(dev), 0.05, 0.45, top_k, pri.to(dev), cfg(n_classes=Cn))
        got2 = DT.detect_objects(locs.to(dev), scores.to(dev), 0.05, 0.45, top_k, pri.to(dev))
        for i in range(N):
            assert want[1][i].numel() > 5
            for g in (got, got2):
                assert torch.equal(g[1][i].cpu(), want[1][i]), (table, i)
                assert torch.allclose(g[2][i].cpu(), want[2][i], rtol=1e-5, atol=1e-8)
                assert torch.allclose(g[0][i].cpu(), want[0][i], rtol=1e-5, atol=1e-6)
    # sigmoid scores, nothing above the threshold -> placeholder
    pri = PR.ssd300_priors()
    locs, scores = synth.make_eval_batch(pri, 2, 7, 5, bg_bias=0.0)
    want = O.detect_objects(locs.clone(), scores - 12.0, 0.5, 0.45, 200, pri, focal_type="sigmoid")
    got = MU.detect_objects(locs.to(dev), (scores - 12.0).to(dev), 0.5, 0.45, 200, pri.to(dev),
                            cfg(n_classes=7, focal_type="sigmoid"))
    for i in range(2):
        assert got[1][i].tolist() == want[1][i].tolist() == [0]
        assert got[0][i].cpu().tolist() == [[0.0, 0.0, 1.0, 1.0]]


def test_second_band_when_candidates_run_out(dev):
    """Config 2 shape, but every prior of two images decodes to one of a handful of boxes: NMS keeps a few boxes
    per class, the first band of candidates (those above the per-image cutoff) runs out before top_k + 1 boxes
    survive, and the NMS kernel must fetch the rest itself (second band) - exactly the reference's answer,
    class-major because fewer than top_k boxes survive. A normal image sits between them."""
    import torchvision
    import shape_based_object_detection_b200 as S
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    pri = PR.ssd512_canonical_priors()
    P = pri.size(0)
    _, scores = synth.make_eval_batch(pri, 3, 81, 4321 + 9)
    gen = torch.Generator().manual_seed(3)
    base = torch.tensor([[0.10, 0.10, 0.30, 0.35], [0.55, 0.15, 0.90, 0.50]])  # two boxes: <= 160 survivors
    locs = base[torch.randint(0, 2, (3, P), generator=gen)].contiguous()
    locs[1] = torch.rand((P, 4), generator=gen) * 0.5
    locs[1, :, 2:] += locs[1, :, :2] + 0.02  # a normal image: many distinct boxes
    want = O.detect(locs.clone(), scores, 0.01, 0.45, 200, pri, box_type="corner", nms_fn=torchvision.ops.nms,
                    return_priors=True)
    for _ in range(2):  # twice: the workspace is left clean
        got = S.detect_batched(locs.to(dev).contiguous(), scores.to(dev), 0.01, 0.45, 200, None, box_type="corner",
                               clamp_inplace=True)
        assert not _check_detect(got, want, 3)
    assert want[1][0].numel() < 200 and want[1][2].numel() < 200 and want[1][1].numel() == 200


def test_detect_tools_kept_list_spills_to_global_memory(dev):
    """detect_tools: the FIRST (per-class) NMS stage may keep far more boxes than the second, class-agnostic one
    lets through. 150 well-separated boxes x 80 classes, every class above min_score: 12 000 first-stage
    survivors (more than the 5 120 that fit in shared memory) and 150 final boxes. The reference has no limit
    here; the kept list spills to global memory."""
    import torchvision
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200.detect_scripts import detect_tools as DT
    gen = torch.Generator().manual_seed(11)
    n_pos, Cn = 150, 81
    gx, gy = torch.meshgrid(torch.arange(15.), torch.arange(10.), indexing="ij")
    x0, y0 = gx.reshape(-1) / 15.0, gy.reshape(-1) / 10.0
    boxes = torch.stack([x0 + 0.005, y0 + 0.005, x0 + 0.055, y0 + 0.085], 1)[None].contiguous()  # [1,150,4]
    scores = torch.randn((1, n_pos, Cn), generator=gen) * 0.05  # softmax ~ 1/81 = 0.0123 > 0.01 for (almost) all
    want = O.detect(boxes.clone(), scores, 0.01, 0.45, 200, None, box_type="corner", second_nms=0.7,
                    nms_fn=torchvision.ops.nms)
    got = DT.detect_refine(boxes.to(dev).contiguous(), scores.to(dev), 0.01, 0.45, 200, None)
    assert want[1][0].numel() == n_pos
    assert torch.equal(got[1][0].cpu(), want[1][0])
    assert torch.allclose(got[2][0].cpu(), want[2][0], rtol=1e-5, atol=1e-8)
    assert torch.allclose(got[0][0].cpu(), want[0][0], rtol=1e-5, atol=1e-6)


def test_config2_detect_indices_exact_given_score_bits(dev):
    """Policy for index exactness: the kept (class, prior) sequence is bit-exact GIVEN the
    score bits. The bits themselves come from an activation computed as torch does (max shift, accurate exp,
    true division) but with another summation order, so they agree with a CPU run to a few ulp - enough for two
    same-class candidates an ulp apart to swap places on rare images. Checked separately on all 32 images of
    config 2: (1) probabilities within 4e-6 relative of torch's; (2) the oracle run ON THE GPU's PROBABILITIES
    reproduces the kept sequence exactly, priors included."""
    import ctypes as C
    import torchvision
    import shape_based_object_detection_b200 as S
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import _lib as L
    from shape_based_object_detection_b200 import priors as PR, synth
    pri = PR.ssd512_canonical_priors()
    locs, scores = synth.make_eval_batch(pri, 32, 81, 4321)
    s_d = scores.to(dev)
    probs = torch.empty_like(s_d)
    L.check(L.lib().sbod_detect_probabilities(L.ptr(s_d), 32, pri.size(0), 81, L.ACT_SOFTMAX, L.ptr(probs), L.stream_ptr()))
    probs = probs.cpu()
    ref = scores.softmax(dim=2)
    rel = ((probs - ref).abs() / ref.clamp_min(1e-30)).max().item()
    assert rel <= 4e-6, rel
    want = O.detect(locs.clone(), probs, 0.01, 0.45, 200, pri, focal_type="none_is_identity",
                    nms_fn=torchvision.ops.nms, return_priors=True)
    got = S.detect_batched(locs.to(dev), s_d, 0.01, 0.45, 200, pri.to(dev))
    bad = _check_detect(got, want, 32)
    assert not bad, "kept sequence differs from the oracle run on identical score bits on images %s" % bad


def test_prior_tables_generated_on_the_gpu(dev):
    """the prior / anchor / location generators (models/SSD300.py:389-443 etc., python triple
    loops at model construction) as one kernel: bit-identical to the host tables, which the CPU suite pins to
    the reference's own generators."""
    from shape_based_object_detection_b200 import priors as PR
    from shape_based_object_detection_b200.models import compute_location
    for name, fn in PR.PRIOR_TABLES.items():
        host, gpu = fn(), fn(device=dev)
        assert gpu.is_cuda and gpu.shape == host.shape, name
        assert torch.equal(gpu.cpu(), host), name
    for size in (None, (800, 1333)):
        host = compute_location(image_size=size)
        gpu = compute_location(device=dev, image_size=size)
        assert len(host) == len(gpu) == 5
        assert all(torch.equal(a, b.cpu()) for a, b in zip(host, gpu))


def test_packed_ground_truth_single_copy(dev):
    """the batch's ground truth in one pinned CSR buffer (dataset.collate), ONE H2D copy; the
    loss modules take it in place of the `boxes` list and give the same loss and gradients bit for bit."""
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.dataset.collate import collate_fn
    from shape_based_object_detection_b200.models import MultiBoxLoss512
    pri = PR.ssd512_priors()
    locs, scores, bx, lb = synth.make_train_batch(pri, 6, 21, 30, 9)
    batch = [(torch.zeros(1), bx[i], lb[i], i, torch.zeros(bx[i].size(0), dtype=torch.uint8)) for i in range(6)]
    _, boxes, labels, _, _ = collate_fn(batch)
    assert boxes.packed.buf.is_pinned()
    crit = MultiBoxLoss512(pri.to(dev), cfg(n_classes=21))
    outs = []
    for mode in ("lists_on_device", "lists_on_host", "packed"):
        l_d, s_d = locs.to(dev).requires_grad_(True), scores.to(dev).requires_grad_(True)
        if mode == "lists_on_device":
            loss = crit(l_d, s_d, [b.to(dev) for b in boxes], [l.to(dev) for l in labels])
        elif mode == "lists_on_host":
            loss = crit(l_d, s_d, list(boxes), labels)
        else:
            loss = crit(l_d, s_d, boxes.packed.to(dev), None)
        loss.backward()
        outs.append((loss.item(), l_d.grad.clone(), s_d.grad.clone()))
    for o in outs[1:]:
        assert o[0] == outs[0][0] and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])


def test_random_crop_matches_the_reference_and_its_random_stream(golden, dev):
    """dataset/transforms.py:124-205 random_crop, the CPU user of find_jaccard_overlap, on CUDA
    tensors: the overlaps of a round's trial crops come from one dense-IoU launch. Fixtures from the reference
    itself (oracle/make_golden.py --only-crop): crop, boxes, labels and the NEXT draw of python's random."""
    import random
    from cases import crop_inputs
    from shape_based_object_detection_b200.dataset.transforms import random_crop
    G = golden["crop"]
    cropped = 0
    for seed in range(24):
        image, boxes, labels = crop_inputs(seed)
        random.seed(1000 + seed)
        ni, nb, nl = random_crop(image.to(dev), boxes.to(dev), labels.to(dev))
        assert tuple(ni.shape) == tuple(G["s%d_shape" % seed].tolist()), seed
        assert abs(ni.double().sum().item() - float(G["s%d_sum" % seed])) <= 1e-6 * float(G["s%d_sum" % seed])
        assert torch.equal(nb.cpu(), torch.from_numpy(G["s%d_boxes" % seed])), seed
        assert torch.equal(nl.cpu(), torch.from_numpy(G["s%d_labels" % seed])), seed
        assert random.random() == float(G["s%d_next" % seed]), seed
        cropped += int(tuple(ni.shape) != tuple(image.shape))
    assert cropped >= 10


def test_hard_negative_mining_with_massive_ties(dev):
    """Hard-negative mining when (nearly) every candidate has the same cross entropy (constant logits): the k-th
    candidate's leading digit then holds more values than the mining tail keeps in shared memory (its multi-pass
    path), and only some of the ties at the threshold belong to the top-k (taken by ascending prior index, like
    a stable sort). Loss against the oracle; the gradient touches exactly n_pos + 3 n_pos rows per image."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512, RetinaFocalLoss
    pri = PR.ssd512_canonical_priors()
    N, Cn = 2, 5
    locs, scores, bx, lb = synth.make_train_batch(pri, N, Cn, 20, 31)
    for variant, Mod, scale in (("s512", MultiBoxLoss512, 0.0), ("ret", RetinaFocalLoss, 0.0), ("s512", MultiBoxLoss512, 1e-3)):
        sc = scores * scale  # 0: all CEs equal ln(5); 1e-3: all inside one leading radix digit, (almost) no exact ties
        want, parts = O.multibox_loss(variant, pri, locs, sc, bx, lb, want_parts=True)
        crit = Mod(pri.to(dev), cfg(n_classes=Cn))
        l_d, s_d = locs.to(dev).requires_grad_(True), sc.to(dev).requires_grad_(True)
        loss = crit(l_d, s_d, [b.to(dev) for b in bx], [l.to(dev) for l in lb])
        loss.backward()
        assert abs(loss.item() - want.item()) <= RTOL_LOSS * abs(want.item()), (variant, scale, loss.item(), want.item())
        rows = (s_d.grad.abs().sum(2) > 0).sum(1).cpu()
        assert torch.equal(rows, 4 * parts["n_pos"]), (rows, parts["n_pos"])


def test_hard_negative_mining_outside_the_histogram_range(dev):
    """The mining histogram resolves the 16 octaves [2^-10, 2^6); everything smaller shares bin 0 and everything
    larger bin 4095, where the k-th value is found by full 32-bit radix passes. Case "tiny": a confident model
    (background logit + 12): nearly every candidate CE is below 2^-10, so the threshold lies in bin 0 (more values
    than the shared list holds: the passes run over global memory). Case "huge": logits x 60: CEs of several
    hundred, the threshold lies in bin 4095. Case "mixed": both tails populated, threshold in a regular bin.
    Loss, gradients and the number of gradient rows against the oracle."""
    from oracle import box_pipeline as O
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512
    pri = PR.ssd512_canonical_priors()
    N, Cn = 2, 7
    locs, scores, bx, lb = synth.make_train_batch(pri, N, Cn, 30, 32)
    cases = {}
    tiny = scores.clone()
    tiny[:, :, 0] += 12.0
    cases["tiny"] = tiny
    cases["huge"] = scores * 60.0
    mixed = scores.clone()
    mixed[:, ::3, 0] += 14.0
    mixed[:, 1::3, :] *= 50.0
    cases["mixed"] = mixed
    for name, sc in cases.items():
        l_c, s_c = locs.clone().requires_grad_(True), sc.clone().requires_grad_(True)
        want, parts = O.multibox_loss("s512", pri, l_c, s_c, bx, lb, want_parts=True)
        want.backward()
        crit = MultiBoxLoss512(pri.to(dev), cfg(n_classes=Cn))
        l_d, s_d = locs.to(dev).requires_grad_(True), sc.to(dev).requires_grad_(True)
        loss = crit(l_d, s_d, [b.to(dev) for b in bx], [l.to(dev) for l in lb])
        loss.backward()
        assert abs(loss.item() - want.item()) <= RTOL_LOSS * abs(want.item()), (name, loss.item(), want.item())
        assert torch.allclose(l_d.grad.cpu(), l_c.grad, rtol=RTOL_GRAD, atol=1e-7), name
        got_g, want_g = s_d.grad.cpu(), s_c.grad
        got_m, want_m = got_g.abs().sum(2) > 0, want_g.abs().sum(2) > 0
        # (rows whose softmax - onehot underflows to exactly zero in every class do not count on either side)
        assert torch.equal(got_m.sum(1), want_m.sum(1)), (name, got_m.sum(1), want_m.sum(1), parts["n_pos"])
        # Cross entropies of ~1e-4 are only known to ~1e-7 absolute in fp32 (log of 1 + tiny) on either side, so two
        # candidates that close at the threshold may be taken the other way round: at most one swap per image,
        # between rows with the same gradient to 1e-3; every other row to the usual tolerance.
        both = got_m & want_m
        assert int((got_m ^ want_m).sum()) <= 2 * N, (name, int((got_m ^ want_m).sum()))
        for n in range(N):
            only_g, only_w = got_g[n][got_m[n] & ~want_m[n]], want_g[n][want_m[n] & ~got_m[n]]
            assert only_g.shape == only_w.shape
            if only_g.numel():
                assert torch.allclose(only_g[:, 0].sort().values, only_w[:, 0].sort().values, rtol=1e-2, atol=1e-9), name
        assert torch.allclose(got_g[both], want_g[both], rtol=RTOL_GRAD, atol=1e-7), name


def test_randomised_shapes_against_the_oracle():
    """tools/stress.py on a few random cases: odd prior counts, class counts on both sides of the compile-time
    specialisations (2 ... 130), zero to 300 objects per image, the four criteria, scaled logits; train path
    (assignment bit-exact, loss 1e-5, gradients 1e-4) and eval path (labels exact, scores / boxes 1e-5)."""
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "tools", "stress.py"), "8", "7"], capture_output=True,
                         text=True, timeout=600, cwd=repo)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-2000:])
    assert "stress: 0 mismatches in 8 cases" in out.stdout
