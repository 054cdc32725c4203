#!/usr/bin/env python
"""Per-kernel steady-state timings (CUDA events, hot clocks) of the bench workload. GPU box only."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shape_based_object_detection_b200 as S  # noqa: E402
from shape_based_object_detection_b200 import _lib as L, core, priors as PR, synth  # noqa: E402

NAME = sys.argv[1] if len(sys.argv) > 1 else "ssd512_canonical"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
Cn = int(sys.argv[3]) if len(sys.argv) > 3 else 81
G = int(sys.argv[4]) if len(sys.argv) > 4 else 100
REPS = 30
dev = torch.device("cuda:0")
pri = PR.PRIOR_TABLES[NAME]()
P = pri.size(0)
locs, scores, bx, lb = synth.make_train_batch(pri, N, Cn, G, 1236)
elocs, escores = synth.make_eval_batch(pri, N, Cn, 4321)
pri_d = pri.to(dev)
from shape_based_object_detection_b200.dataset.transforms import cxcy_to_xy  # noqa: E402
pxy = cxcy_to_xy(pri_d)
gt = core.pack_ground_truth([b.to(dev) for b in bx], [l.to(dev) for l in lb], dev)
REG = int(os.environ.get("KB_REG", L.REG_SMOOTH_L1))
CLS = int(os.environ.get("KB_CLS", L.CLS_CE_MINE_NONPOS))
spec = core.LossSpec(reg_kind=REG, cls_kind=CLS)
st = core.LossState(spec, pri_d, pxy, locs.to(dev), scores.to(dev), gt, prefill_grad=bool(int(os.environ.get("KB_PREFILL", "0"))))
lib = L.lib()
sp = L.stream_ptr()
for key, env in ((0, "KB_PDL"),):  # sbod_set_option switches for A/B runs
    if env in os.environ:
        L.check(lib.sbod_set_option(key, int(os.environ[env])))


def ev():
    return torch.cuda.Event(enable_timing=True)


def time_seq(fns, reps=REPS):
    """fns: list of callables run back to back each rep; returns avg ms of each."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    acc = [0.0] * len(fns)
    for _ in range(reps):
        marks = [ev() for _ in range(len(fns) + 1)]
        marks[0].record()
        for i, f in enumerate(fns):
            f()
            marks[i + 1].record()
        torch.cuda.synchronize()
        for i in range(len(fns)):
            acc[i] += marks[i].elapsed_time(marks[i + 1])
    return [a / reps for a in acc]


st.forward()
torch.cuda.synchronize()
print("loss", st.loss.tolist())
t = time_seq([lambda: L.check(lib.sbod_loss_forward_stage(C.byref(st.desc), 0, sp)),
              lambda: L.check(lib.sbod_loss_forward_stage(C.byref(st.desc), 1, sp))])
gl = torch.ones(1, device=dev)
g_l = torch.empty_like(st.locs)
g_s = torch.empty_like(st.scores)
tb = time_seq([lambda: L.check(lib.sbod_loss_backward(C.byref(st.desc), L.ptr(gl), L.ptr(g_l), L.ptr(g_s), sp))])
tf = time_seq([lambda: st.forward()])
# detect
el, es = elocs.to(dev), escores.to(dev)
out = S.detect_batched(el, es, 0.01, 0.45, 200, pri_d)
torch.cuda.synchronize()
cap = 200
ob = torch.empty((N, cap, 4), device=dev)
ol = torch.empty((N, cap), dtype=torch.int64, device=dev)
osc = torch.empty((N, cap), device=dev)
op = torch.empty((N, cap), dtype=torch.int32, device=dev)
oc = torch.empty((N,), dtype=torch.int32, device=dev)
d = L.DetectDesc()
d.locs, d.scores, d.priors_cxcy = el.data_ptr(), es.data_ptr(), pri_d.data_ptr()
d.N, d.P, d.C = N, P, Cn
d.act_kind, d.box_kind, d.clamp_inplace = 0, 0, 0
d.min_score, d.max_overlap, d.top_k, d.second_nms_thr, d.pre_nms_topk = 0.01, 0.45, 200, -1.0, 0
d.class_agnostic = 0
d.out_boxes, d.out_labels, d.out_scores, d.out_prior, d.out_counts, d.out_cap = (
    ob.data_ptr(), ol.data_ptr(), osc.data_ptr(), op.data_ptr(), oc.data_ptr(), cap)
nb = lib.sbod_detect_workspace_bytes(C.byref(d))
ws = L.Workspace.get(dev, "detect", nb, zero_bytes=lib.sbod_detect_workspace_zero_bytes(C.byref(d)), layout=(N, Cn))
d.workspace, d.workspace_bytes = ws.data_ptr(), nb
td = time_seq([lambda: L.check(lib.sbod_detect_stage(C.byref(d), 2, sp)),
               lambda: L.check(lib.sbod_detect_stage(C.byref(d), 3, sp)),
               lambda: L.check(lib.sbod_detect_stage(C.byref(d), 1, sp))])
byt = N * P * Cn * 4
print(f"shape {NAME} N={N} P={P} C={Cn} G<={G}  logits {byt/1e6:.1f} MB")
for nm, ms in (("match_lse", t[0]), ("classify_mine", t[1]), ("fwd (2 kernels, python call)", tf[0]),
               ("loss_bwd", tb[0]), ("detect_bound (streaming)", td[0]), ("detect_refine", td[1]),
               ("detect_nms", td[2])):
    print(f"{nm:32s} {ms*1e3:9.1f} us   {byt/ms/1e6:8.1f} GB/s of logits")
