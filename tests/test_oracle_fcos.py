"""CPU suite: the FCOS restatement in the oracle (parity unpinned with respect to the reference, whose
FCOS code does not run) is cross-checked against a plain-Python loop statement of the same rules on
a sample of locations, so that the vectorised oracle the CUDA kernels are tested against is itself
checked by an independent formulation."""
import math

import torch

from oracle import box_pipeline as O
from shape_based_object_detection_b200 import synth


def brute_force_target(x, y, level, boxes, labels, center_sample):
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    rad = float(f32(O.FCOS_STRIDES[level]) * f32(O.FCOS_RADIUS))
    lo, hi = float(f32(O.FCOS_SIZES[level][0])), float(f32(O.FCOS_SIZES[level][1]))
    best, best_area, best_t = None, O.FCOS_INF, None
    for g in range(boxes.size(0)):
        x1, y1, x2, y2 = [boxes[g, k] for k in range(4)]
        l, t, r, b = x - x1, y - y1, x2 - x, y2 - y
        if center_sample:
            cx, cy = (x1 + x2) / 2., (y1 + y2) / 2.
            bx0 = cx - rad if cx - rad > x1 else x1
            by0 = cy - rad if cy - rad > y1 else y1
            bx1 = cx + rad if cx + rad < x2 else x2
            by1 = cy + rad if cy + rad < y2 else y2
            inside = min(x - bx0, y - by0, bx1 - x, by1 - y) > 0
        else:
            inside = min(l, t, r, b) > 0
        mx = max(l, t, r, b)
        if inside and lo <= mx <= hi:
            area = (x2 - x1) * (y2 - y1)
            if area < best_area:
                best, best_area, best_t = g, area, (l, t, r, b)
    if best is None:
        return 0, None
    return int(labels[best]), best_t


def test_fcos_assignment_against_python_loops():
    locations = O.fcos_locations()
    allp = torch.cat(locations, 0)
    level_of = torch.cat([torch.full((l.size(0),), k) for k, l in enumerate(locations)])
    gen = torch.Generator().manual_seed(77)
    bx, lb = synth.make_gt(1, 14, 9, gen)
    boxes = torch.cat([bx[0], torch.tensor([[0.05, 0.05, 0.95, 0.95], [0.3, 0.3, 0.34, 0.33]])])
    labels = torch.cat([lb[0], torch.tensor([2, 7])])
    for cs in (True, False):
        lab, tgt = O.fcos_assign(locations, boxes, labels, center_sample=cs)
        assert int((lab > 0).sum()) > 10
        pick = torch.cat([(lab > 0).nonzero().squeeze(1)[:150],
                          torch.randperm(allp.size(0), generator=gen)[:150]])
        for p in pick.tolist():
            want_lab, want_t = brute_force_target(allp[p, 0], allp[p, 1], int(level_of[p]), boxes, labels, cs)
            assert int(lab[p]) == want_lab, p
            if want_lab:
                assert all(float(a) == float(b) for a, b in zip(tgt[p], want_t)), p


def test_fcos_loss_composition():
    """loss = focal/(n_pos+N) + alpha*weighted DIoU + BCE; zero objects in range -> only the focal term's 0."""
    locations = O.fcos_locations()
    P = sum(l.size(0) for l in locations)
    gen = torch.Generator().manual_seed(3)
    bx, lb = synth.make_gt(2, 6, 5, gen)
    locs = torch.rand((2, P, 4), generator=gen) * 0.2 + 0.01
    scores = torch.randn((2, P, 5), generator=gen)
    ctr = torch.randn((2, P), generator=gen)
    total, parts = O.fcos_loss(locations, locs, scores, ctr, bx, lb, alpha=2.0, want_parts=True)
    assert math.isclose(float(total), float(parts["conf"] + 2.0 * parts["loc"] + parts["center"]), rel_tol=1e-6)
    tiny = [torch.tensor([[0.5, 0.5, 0.5005, 0.5005]])] * 2  # too small to contain any cell centre
    total0, parts0 = O.fcos_loss(locations, locs, scores, ctr, tiny, [torch.tensor([1])] * 2, want_parts=True)
    assert int((parts0["labels"] > 0).sum()) == 0 and float(total0) == 0.0
