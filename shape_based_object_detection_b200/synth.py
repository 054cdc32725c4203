"""Seeded synthetic batches of the shapes BASELINE.json names. Generated on the CPU
with torch.Generator so the CUDA path, the oracle and the golden fixtures see identical inputs."""
import torch


def make_gt(n_images, gmax, n_classes, gen, dense=False, gmin=1):
    """list[N] of xyxy boxes [G_i,4] in [0,1] and labels [G_i] in 1..C-1; G_i ~ U{gmin..gmax}."""
    boxes, labels = [], []
    for _ in range(n_images):
        g = int(torch.randint(gmin, gmax + 1, (1,), generator=gen))
        c = torch.rand((g, 2), generator=gen) * 0.8 + 0.1
        hi = 0.12 if dense else 0.32
        wh = torch.rand((g, 2), generator=gen) * (hi - 0.02) + 0.02
        b = torch.cat([c - wh / 2, c + wh / 2], 1).clamp_(0, 1)
        boxes.append(b)
        labels.append(torch.randint(1, n_classes, (g,), generator=gen, dtype=torch.int64))
    return boxes, labels


def make_train_batch(priors_cxcy, n_images, n_classes, gmax, seed, dense=False):
    gen = torch.Generator().manual_seed(seed)
    P = priors_cxcy.size(0)
    boxes, labels = make_gt(n_images, gmax, n_classes, gen, dense=dense)
    locs = torch.randn((n_images, P, 4), generator=gen) * 0.1
    scores = torch.randn((n_images, P, n_classes), generator=gen)
    return locs, scores, boxes, labels


def make_eval_batch(priors_cxcy, n_images, n_classes, seed, bg_bias=8.0, loc_std=0.5, logit_std=2.0):
    """Eval logits ~ N(0, 2^2) with the background logit raised by 8: about 6 % of (prior, class)
    scores pass min_score = 0.01."""
    gen = torch.Generator().manual_seed(seed)
    P = priors_cxcy.size(0)
    locs = torch.randn((n_images, P, 4), generator=gen) * loc_std
    scores = torch.randn((n_images, P, n_classes), generator=gen) * logit_std
    scores[:, :, 0] += bg_bias
    return locs, scores


def adversarial_gt(priors_xy, n_classes):
    """One image exercising the assignment corner cases: duplicate GT, GT with zero overlap
    to every prior, degenerate (w=h=0) GT, two GT sharing a best prior, GT exactly equal to a prior."""
    p = priors_xy
    mid = p[p.size(0) // 2].clone()
    far = torch.tensor([0.0, 0.0, 1e-4, 1e-4])            # overlaps nothing meaningfully
    degenerate = torch.tensor([0.5, 0.5, 0.5, 0.5])        # zero-size GT -> masked to 0
    tiny_a = torch.tensor([0.30, 0.30, 0.31, 0.31])
    tiny_b = torch.tensor([0.302, 0.302, 0.312, 0.312])    # shares its best prior with tiny_a (likely)
    normal = torch.tensor([0.55, 0.20, 0.85, 0.60])
    boxes = torch.stack([degenerate, normal, normal.clone(), mid, far, tiny_a, tiny_b]).clamp_(0, 1)
    labels = (torch.arange(boxes.size(0)) % (n_classes - 1) + 1).to(torch.int64)
    return boxes, labels


def make_map_case(n_images, n_classes, gmax, dets_per_image, seed, p_difficult=0.15, ties=False):
    """Ground truth + detections for metrics.calculate_mAP: a share of the detections are jittered
    copies of objects (some twice, so that duplicates and second-best matches occur), the rest random
    boxes; the last class has no detections. ties=True adds exact score ties (their order is
    unspecified in the reference, whose sort is unstable; the oracle and the kernels take them in the
    order of the concatenated lists)."""
    gen = torch.Generator().manual_seed(seed)
    true_boxes, true_labels = make_gt(n_images, gmax, n_classes - 1, gen)   # labels 1..n_classes-2: last class unused
    true_diff = [(torch.rand((b.size(0),), generator=gen) < p_difficult).to(torch.uint8) for b in true_boxes]
    det_boxes, det_labels, det_scores = [], [], []
    for i in range(n_images):
        g = true_boxes[i].size(0)
        k = int(torch.randint(0, dets_per_image + 1, (1,), generator=gen))
        src = torch.randint(0, g, (k,), generator=gen)
        near = torch.rand((k,), generator=gen) < 0.6
        jit = (torch.rand((k, 4), generator=gen) - 0.5) * 0.08
        b = torch.where(near[:, None], true_boxes[i][src] + jit, torch.rand((k, 4), generator=gen))
        b = torch.cat([torch.minimum(b[:, :2], b[:, 2:]), torch.maximum(b[:, :2], b[:, 2:])], 1).clamp_(0, 1)
        lab = torch.where(torch.rand((k,), generator=gen) < 0.8, true_labels[i][src],
                          torch.randint(1, n_classes - 1, (k,), generator=gen))
        sc = torch.rand((k,), generator=gen)
        if ties and k >= 4:
            sc[1] = sc[0]  # exact ties
            sc[3] = sc[2]
        det_boxes.append(b)
        det_labels.append(lab.to(torch.int64))
        det_scores.append(sc)
    return det_boxes, det_labels, det_scores, true_boxes, true_labels, true_diff
