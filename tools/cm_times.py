#!/usr/bin/env python
"""Phase time stamps of classify_kernel / mine_kernel (needs a libsbod.so built with -DSBOD_DEBUG_HOOKS). GPU box only."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shape_based_object_detection_b200 import _lib as L, core, priors as PR, synth  # noqa: E402
from shape_based_object_detection_b200.dataset.transforms import cxcy_to_xy  # noqa: E402

NAME = sys.argv[1] if len(sys.argv) > 1 else "ssd512_canonical"
N, Cn, G = 32, int(sys.argv[2]) if len(sys.argv) > 2 else 81, int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda:0")
pri = PR.PRIOR_TABLES[NAME]()
locs, scores, bx, lb = synth.make_train_batch(pri, N, Cn, G, 1236)
pri_d = pri.to(dev)
gt = core.pack_ground_truth([b.to(dev) for b in bx], [l.to(dev) for l in lb], dev)
spec = core.LossSpec(reg_kind=int(os.environ.get("KB_REG", 1)), cls_kind=int(os.environ.get("KB_CLS", 0)))
st = core.LossState(spec, pri_d, cxcy_to_xy(pri_d), locs.to(dev), scores.to(dev), gt, prefill_grad=True)
lib = L.lib()
sp = L.stream_ptr()
buf = (C.c_ulonglong * (8 + 4 * 64))()
for it in range(3):
    L.check(lib.sbod_loss_forward_stage(C.byref(st.desc), 0, sp))
    torch.cuda.synchronize()
    lib.sbod_debug_cm_times(None, 1)
    L.check(lib.sbod_loss_forward_stage(C.byref(st.desc), 1, sp))
    torch.cuda.synchronize()
    lib.sbod_debug_cm_times(buf, 0)
    t = np.array(list(buf), dtype=np.float64)
    t0 = t[0]
    rel = lambda x: (x - t0) / 1e3
    img = t[8:8 + 4 * N].reshape(N, 4)
    print(f"iter {it}: wait-done(min) {rel(t[1]):.1f}  forced(max) {rel(t[2]):.1f}  phaseA(max) {rel(t[3]):.1f} "
          f"phaseB(max) {rel(t[4]):.1f}  classify end(max) {rel(t[5]):.1f}  end(max) {rel(t[7]):.1f} us")
    print("   mine start      ", np.round(rel(img[:, 0]), 1).tolist())
    print("   bin found (+us) ", np.round((img[:, 1] - img[:, 0]) / 1e3, 1).tolist())
    print("   scan done (+)   ", np.round((img[:, 2] - img[:, 0]) / 1e3, 1).tolist())
    print("   batch ticket (+)", np.round((img[:, 3] - img[:, 0]) / 1e3, 1).tolist())
