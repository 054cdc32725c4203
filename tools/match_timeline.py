#!/usr/bin/env python
"""Per-warp timeline of the match role (SBOD_DEBUG_SKIP must have bit 2 set). GPU box only."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shape_based_object_detection_b200 import _lib as L, core, priors as PR, synth  # noqa: E402
from shape_based_object_detection_b200.dataset.transforms import cxcy_to_xy  # noqa: E402

N, Cn = 32, 81
G = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
pri = PR.PRIOR_TABLES["ssd512_canonical"]()
locs, scores, bx, lb = synth.make_train_batch(pri, N, Cn, G, 1236)
pri_d = pri.to(dev)
gt = core.pack_ground_truth([b.to(dev) for b in bx], [l.to(dev) for l in lb], dev)
st = core.LossState(core.LossSpec(reg_kind=L.REG_SMOOTH_L1, cls_kind=L.CLS_CE_MINE_NONPOS), pri_d, cxcy_to_xy(pri_d), locs.to(dev), scores.to(dev), gt)
lib, sp = L.lib(), L.stream_ptr()
for _ in range(5):
    L.check(lib.sbod_loss_forward_stage(C.byref(st.desc), 0, sp))
torch.cuda.synchronize()
gmax = st.desc.gmax
al = lambda x: (x + 255) // 256 * 256
P = pri.size(0)
off = 256 + al(N * max(gmax, 1) * 8) + al(N * 4) + al(N * P * 8)
nw = 148 * 2 * 8
raw = st.ws[off:off + nw * 64].cpu().numpy().view(np.int64).reshape(nw, 8)
t0, t2 = raw[:, 0], raw[:, 1]
sl, sw = raw[:, 2] & 0xffffffff, raw[:, 2] >> 32
ph = raw[:, 3:8]
base = t0.min()
print("warps", nw, "G<=", G, " (library must be built with -DSBOD_MATCH_TIMELINE)")
print("start  min/med/max us", (t0.min() - base) / 1e3, np.median(t0 - base) / 1e3, (t0.max() - base) / 1e3)
print("end    min/med/max us", (t2.min() - base) / 1e3, np.median(t2 - base) / 1e3, (t2.max() - base) / 1e3)
print("items  min/med/max", sl.min(), np.median(sl), sl.max(), " image switches med/max", np.median(sw), sw.max())
names = ["get item", "priors", "bbox", "groups", "outputs"]
tot = ph.sum()
for i, nm in enumerate(names):
    print(f"phase {nm:10s} mean cycles/warp {ph[:, i].mean():9.0f}  per item {ph[:, i].sum() / sl.sum():7.0f}  share {100.0 * ph[:, i].sum() / tot:5.1f}%")
h, e = np.histogram((t2 - base) / 1e3, bins=10)
print("end-time histogram", list(zip(np.round(e[:-1], 1).tolist(), h.tolist())))
