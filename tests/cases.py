"""Shared definitions of the small parity cases: used by oracle/make_golden.py (reference run) and by
the tests (oracle and CUDA runs) so that all three see identical seeded inputs."""
from shape_based_object_detection_b200 import priors as PR


def case_priors(case):
    return PR.PRIOR_TABLES[case["priors"]]()[:: case["stride"]].contiguous()


LOSS_CASES = {
    "s300_l1_ce": dict(variant="s300", priors="ssd300", stride=5, N=2, C=6, gmax=6, seed=11, reg="", cls=""),
    "s300_diou_focal": dict(variant="s300", priors="ssd300", stride=5, N=2, C=6, gmax=6, seed=12, reg="DIOU",
                            cls="FOCAL"),
    "s512_sl1_ce": dict(variant="s512", priors="ssd512", stride=5, N=2, C=6, gmax=8, seed=13, reg="", cls="",
                        adversarial=True),
    "s512_diou_focal": dict(variant="s512", priors="ssd512", stride=5, N=2, C=6, gmax=8, seed=14, reg="DIOU",
                            cls="FOCAL"),
    "s512_thr06": dict(variant="s512", priors="ssd512", stride=5, N=2, C=5, gmax=8, seed=15, reg="", cls="",
                       threshold=0.6),
    "ret_sl1_ce": dict(variant="ret", priors="retinanet", stride=9, N=2, C=6, gmax=8, seed=16, reg="", cls=""),
    "ret_diou_focal": dict(variant="ret", priors="retinanet", stride=9, N=2, C=6, gmax=8, seed=17, reg="DIOU",
                           cls="FOCAL"),
    "rfd": dict(variant="rfd", priors="refinedet512", stride=6, N=2, C=4, gmax=12, seed=18),
}

_D = dict(priors="ssd300", stride=3, N=2, C=6, seed=21, bg=4.0, min_score=0.01, max_overlap=0.45, top_k=200)
DETECT_CASES = {
    "det_base": dict(_D, fn="utils.detect"),
    "det_topk20": dict(_D, fn="utils.detect", top_k=20, seed=22),
    "det_few": dict(_D, fn="utils.detect", bg=11.0, seed=23),
    "det_empty": dict(_D, fn="utils.detect", bg=40.0, seed=24),
    "det_sigmoid": dict(_D, fn="utils.detect", focal_type="sigmoid", min_score=0.97, seed=25),
    "det_corner": dict(_D, fn="utils.detect", box_type="corner", seed=26),
    "det_center": dict(_D, fn="utils.detect", box_type="center", seed=27),
    "det_keep": dict(_D, fn="utils.detect", prior_keep=True, seed=28),
    "tools_detect": dict(_D, fn="tools.detect", seed=29),
    "tools_detect_few": dict(_D, fn="tools.detect", bg=10.0, seed=30),
    "tools_refine": dict(_D, fn="tools.refine", prior_keep=True, seed=31),
}


def operator_inputs():
    """Seeded inputs of the stand-alone operator fixtures (tests/golden/operators.npz)."""
    import torch
    from shape_based_object_detection_b200 import synth
    g = torch.Generator().manual_seed(101)
    pri = PR.ssd300_priors()
    pri_xy = torch.cat([pri[:, :2] - pri[:, 2:] / 2, pri[:, :2] + pri[:, 2:] / 2], 1)
    boxes, labels = synth.adversarial_gt(pri_xy, 21)
    sub = pri_xy[::37].clone()
    sub[3] = torch.tensor([0.2, 0.2, 0.2, 0.2])  # zero-size anchor -> -1 column
    m = 200
    b1 = torch.rand((m, 2), generator=g) * 0.6
    b1 = torch.cat([b1, b1 + torch.rand((m, 2), generator=g) * 0.35 + 0.02], 1)
    b2 = torch.rand((m, 2), generator=g) * 0.6
    b2 = torch.cat([b2, b2 + torch.rand((m, 2), generator=g) * 0.35 + 0.02], 1)
    b2[:5] = b1[:5]              # identical boxes (ties in max/min)
    b2[5:10, :2] = b1[5:10, 2:]  # touching corners
    b2[5:10, 2:] = b2[5:10, :2] + 0.1
    ppm = pri[::37][:m].contiguous()
    loc = torch.randn((ppm.size(0), 4), generator=g) * 0.3
    lg = torch.randn((64, 7), generator=g) * 2
    tg = torch.randint(0, 7, (64,), generator=g)
    pr_ = torch.randn((50, 4), generator=g) * 0.2
    tg_ = torch.randn((50, 4), generator=g) * 0.2
    nb = torch.cat([b1, b2], 0)
    ns = torch.rand((nb.size(0),), generator=g)
    return dict(pri=pri, pri_xy=pri_xy, boxes=boxes, labels=labels, sub=sub, b1=b1, b2=b2, ppm=ppm, loc=loc,
                lg=lg, tg=tg, pr=pr_, tgt=tg_, nb=nb, ns=ns, wts=torch.linspace(0.5, 1.5, m))


# metrics.calculate_mAP: synth.make_map_case arguments + IoU threshold
MAP_CASES = {
    "small": dict(n_images=6, n_classes=7, gmax=8, dets=40, seed=5, threshold=0.5),
    "voc_like": dict(n_images=40, n_classes=21, gmax=12, dets=120, seed=6, threshold=0.5),
    "strict": dict(n_images=25, n_classes=11, gmax=20, dets=200, seed=7, threshold=0.75),
    "sparse": dict(n_images=12, n_classes=81, gmax=6, dets=15, seed=8, threshold=0.5),
}


def map_inputs(case):
    from shape_based_object_detection_b200 import synth
    return synth.make_map_case(case["n_images"], case["n_classes"], case["gmax"], case["dets"], case["seed"])


def crop_inputs(seed):
    import torch
    """Seeded image / boxes / labels of the random_crop fixtures (shared with the tests through this module)."""
    g = torch.Generator().manual_seed(seed)
    h, w = 240 + 16 * (seed % 5), 320 - 8 * (seed % 7)
    image = torch.rand((3, h, w), generator=g)
    n = 3 + seed % 6
    c = torch.rand((n, 2), generator=g) * torch.tensor([w * 0.8, h * 0.8]) + torch.tensor([w * 0.1, h * 0.1])
    wh = torch.rand((n, 2), generator=g) * torch.tensor([w * 0.3, h * 0.3]) + 8.0
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    boxes[:, 0::2].clamp_(0, w - 1)
    boxes[:, 1::2].clamp_(0, h - 1)
    labels = torch.randint(1, 21, (n,), generator=g)
    return image, boxes, labels


