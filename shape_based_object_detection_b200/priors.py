"""Prior / anchor box tables (cxcy, clamped to [0,1]) for the detectors on the path.

Table-driven restatement of the reference's python triple loops (models/SSD300.py:389-443,
SSD512.py:417-474, RetinaNet.py:261-300, RefineDet512.py:655-695) — init-time host code, computed in
float64 like the reference's python floats and rounded once to fp32 — plus the BASELINE.json shapes
the reference models do not generate (canonical SSD512 with 24 564 priors, RetinaNet-640 with 76 725)."""
from math import sqrt

import numpy as np
import torch


def _ssd_level_shapes(scales, ratios, k, extra_for_ratio1=True, last_extra=1.0):
    """(w, h) list for one SSD feature map: one box per ratio, plus the geometric-mean box after ratio 1."""
    shapes = []
    s = scales[k]
    for r in ratios[k]:
        shapes.append((s * sqrt(r), s / sqrt(r)))
        if r == 1.0 and extra_for_ratio1:
            extra = sqrt(s * scales[k + 1]) if k + 1 < len(scales) else last_extra
            shapes.append((extra, extra))
    return shapes


def _grid(levels):
    """levels: list of (fmap_dim, [(w,h), ...]); order = level, row i, column j, shape (reference order)."""
    out = []
    for dim, shapes in levels:
        for i in range(dim):
            for j in range(dim):
                cx, cy = (j + 0.5) / dim, (i + 0.5) / dim
                for (w, h) in shapes:
                    out.append((cx, cy, w, h))
    pri = torch.tensor(np.asarray(out, dtype=np.float64).astype(np.float32))
    return pri.clamp_(0, 1)


def ssd300_priors():
    """8 732 priors, SSD300.py:389-443."""
    dims = [38, 19, 10, 5, 3, 1]
    scales = [0.1, 0.2, 0.375, 0.55, 0.725, 0.9]
    ratios = [[1., 2., 0.5], [1., 2., 3., 0.5, .333], [1., 2., 3., 0.5, .333], [1., 2., 3., 0.5, .333],
              [1., 2., 0.5], [1., 2., 0.5]]
    return _grid([(dims[k], _ssd_level_shapes(scales, ratios, k)) for k in range(6)])


def ssd512_priors():
    """10 248 priors of the reference model, SSD512.py:417-474 (no extra box on conv4_3)."""
    dims = [64, 32, 16, 8, 4, 2, 1]
    scales = [0.04, 0.08, 0.16, 0.24, 0.32, 0.64, 0.96]
    ratios = [[1.], [1., 2., 0.5], [1., 2., 3., 0.5, .333], [1., 2., 3., 0.5, .333], [1., 2., 3., 0.5, .333],
              [1., 2., 3., 4., 0.5, .333, 0.25], [1., 2., 3., 4., 0.5, .333, 0.25]]
    return _grid([(dims[k], _ssd_level_shapes(scales, ratios, k, extra_for_ratio1=(k != 0))) for k in range(7)])


def ssd512_canonical_priors():
    """24 564 priors of the canonical SSD512 (BASELINE.json config 2): fmaps (64,32,16,8,4,2,1) with
    (4,6,6,6,6,4,4) boxes per location."""
    dims = [64, 32, 16, 8, 4, 2, 1]
    scales = [0.07, 0.15, 0.30, 0.45, 0.60, 0.75, 0.90]
    r4, r6 = [1., 2., 0.5], [1., 2., 3., 0.5, .333]
    ratios = [r4, r6, r6, r6, r6, r4, r4]
    return _grid([(dims[k], _ssd_level_shapes(scales, ratios, k, last_extra=1.05)) for k in range(7)])


def _retina_levels(dims, scales, factors, ratios=(1., 2., 0.5)):
    levels = []
    for d, s in zip(dims, scales):
        shapes = [(s * f * sqrt(r), s * f / sqrt(r)) for r in ratios for f in factors]
        levels.append((d, shapes))
    return levels


def retinanet_priors():
    """32 736 anchors of the reference model (hard-coded 512 grid), RetinaNet.py:261-300."""
    return _grid(_retina_levels([64, 32, 16, 8, 4], [0.04, 0.08, 0.16, 0.32, 0.64], [2. ** 0, 2. ** (1 / 3.)]))


def retinanet640_priors():
    """76 725 anchors (BASELINE.json config 3): (80,40,20,10,5)^2 x 9 (3 ratios x 3 octave scales)."""
    return _grid(_retina_levels([80, 40, 20, 10, 5], [0.05, 0.1, 0.2, 0.4, 0.8],
                                [2. ** 0, 2. ** (1 / 3.), 2. ** (2 / 3.)]))


def refinedet512_priors():
    """16 320 priors, RefineDet512.py:655-695."""
    return _grid(_retina_levels([64, 32, 16, 8], [0.0625, 0.125, 0.25, 0.5], [1.]))


PRIOR_TABLES = {
    "ssd300": ssd300_priors, "ssd512": ssd512_priors, "ssd512_canonical": ssd512_canonical_priors,
    "retinanet": retinanet_priors, "retinanet640": retinanet640_priors, "refinedet512": refinedet512_priors,
}
