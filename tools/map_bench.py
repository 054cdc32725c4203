#!/usr/bin/env python
"""calculate_mAP on the GPU vs the oracle port on the host (GPU box only).
usage: map_bench.py [n_images] [n_classes] [dets_per_image]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shape_based_object_detection_b200 import synth  # noqa: E402
from shape_based_object_detection_b200.metrics import calculate_mAP  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
Cn = int(sys.argv[2]) if len(sys.argv) > 2 else 81
K = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda:0")
case = synth.make_map_case(N, Cn, 15, K, 3)
label_map = {("background" if i == 0 else "c%d" % i): i for i in range(Cn)}
args = [[t.to(dev) for t in lst] for lst in case]
n_det = sum(int(t.size(0)) for t in case[1])
calculate_mAP(*args, 0.5, label_map)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    aps, m = calculate_mAP(*args, 0.5, label_map)
torch.cuda.synchronize()
t_api = (time.perf_counter() - t0) / 5
print(f"{N} images, {n_det} detections, {Cn} classes: mAP {m:.4f}; public API (list packing + kernels) {t_api * 1e3:.2f} ms")
# oracle port on a bounded sample
from oracle import box_pipeline as O  # noqa: E402
n_s = min(N, 100)
sub = [lst[:n_s] for lst in case]
t0 = time.perf_counter()
O.calculate_mAP(*sub, 0.5, Cn)
t_cpu = time.perf_counter() - t0
print(f"oracle port, {n_s} images on the host: {t_cpu:.2f} s  ->  {t_cpu / n_s * N:.1f} s extrapolated to {N} images")
