#!/usr/bin/env python
"""Phase time stamps of detect_refine_kernel / detect_nms_kernel (needs libsbod.so built with -DSBOD_DEBUG_HOOKS)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shape_based_object_detection_b200 as S  # noqa: E402
from shape_based_object_detection_b200 import _lib as L, core, priors as PR, synth  # noqa: E402

NAME = sys.argv[1] if len(sys.argv) > 1 else "ssd512_canonical"
N, Cn = 32, int(sys.argv[2]) if len(sys.argv) > 2 else 81
dev = torch.device("cuda:0")
pri = PR.PRIOR_TABLES[NAME]()
elocs, escores = synth.make_eval_batch(pri, N, Cn, 4321)
call = core.make_detect_desc(elocs.to(dev), escores.to(dev), 0.01, 0.45, 200, pri.to(dev))
lib = L.lib()
sp = L.stream_ptr()
buf = (C.c_ulonglong * 16)()
for it in range(4):
    L.check(lib.sbod_detect_stage(C.byref(call["desc"]), 2, sp))
    torch.cuda.synchronize()
    lib.sbod_debug_det_times(None, 1)
    L.check(lib.sbod_detect_stage(C.byref(call["desc"]), 3, sp))
    torch.cuda.synchronize()
    lib.sbod_debug_det_times(buf, 0)
    t = np.array(list(buf), dtype=np.float64)
    rel = lambda x: (x - t[0]) / 1e3
    print(f"refine iter {it}: last CTA start {rel(t[1]):.1f}  cutoff(max) {rel(t[2]):.1f}  compaction(max) {rel(t[3]):.1f}  "
          f"eval(max) {rel(t[4]):.1f}  end(max) {rel(t[5]):.1f} us")
    lib.sbod_debug_det_times(None, 1)
    L.check(lib.sbod_detect_stage(C.byref(call["desc"]), 1, sp))
    torch.cuda.synchronize()
    lib.sbod_debug_det_times(buf, 0)
    t = np.array(list(buf), dtype=np.float64)
    print("   nms stamps (us after first CTA start):", [round((x - t[8]) / 1e3, 1) if x else None for x in t[9:16]])
