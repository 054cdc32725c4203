#!/usr/bin/env python
"""NMS phase stamps on the config-4 (RefineDet, two-stage NMS) eval batch. Debug build only."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from shape_based_object_detection_b200 import _lib as L, core
from shape_based_object_detection_b200.models import offset2bbox
dev = torch.device("cuda:0")
pri, train, ev = bench.make_inputs(4, 32, 1238, 4325)
arm_l, odm_l, sc = [t.to(dev) for t in ev["tensors"]]
keep = ev["arm_scores"].to(dev)[:, :, 1] > 0.01
boxes = offset2bbox(arm_l, odm_l, pri.to(dev))
two = int(os.environ.get("TWO", "1"))
call = core.make_detect_desc(boxes, sc, 0.01, 0.45, 200, pri.to(dev), box_type="corner", clamp_inplace=True,
                             prior_keep=keep, second_nms_thr=0.7 if two else -1.0)
lib = L.lib(); sp = L.stream_ptr()
buf = (C.c_ulonglong * 16)()
for it in range(3):
    L.check(lib.sbod_detect_stage(C.byref(call["desc"]), 0, sp))
    torch.cuda.synchronize()
    lib.sbod_debug_det_times(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.sbod_detect_stage(C.byref(call["desc"]), 1, sp))
    e1.record()
    torch.cuda.synchronize()
    lib.sbod_debug_det_times(buf, 0)
    t = np.array(list(buf), dtype=np.float64)
    print("nms ms", e0.elapsed_time(e1), "stamps:", [round((x - t[8]) / 1e3, 1) if x else None for x in t[9:16]], "counts", call["outputs"][4][:8].tolist())
