#!/usr/bin/env python
"""bench.py — images/sec of the detection box pipeline (assign + loss fwd/bwd + NMS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): SSD512 COCO-shaped synthetic batch — 24 564 priors, 81 classes,
<= 100 GT boxes per image, 32 images PER GPU. One step = the train path (targets + loss forward +
backward) on a train batch AND the eval path (decode + threshold + NMS + top-k) on an eval batch of the
same 32 images-per-GPU size. value = images/sec of the whole job with inputs resident in HBM;
e2e = the same through the public Python API with pinned HOST buffers (H2D of every input and D2H
of the results inside the timed region). Multi-GPU: one process per GPU (torchrun), batch sharded
by image, the only exchange is one all-reduce of four loss sums per step ("weak" scaling).
The resident step is replayed as one CUDA graph (--no-graph: eager launches) with the train half and
the eval half on two streams inside it (--no-overlap: one stream); both halves, all nine kernel
launches and the all-reduce are inside the timed region either way.

--impl reference times the CPU oracle port of the reference (oracle/box_pipeline.py — the reference
is pure Python/PyTorch and is not present on the GPU box) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

P_WORKLOAD = "ssd512_canonical"
N_PER_GPU, N_CLASSES, GMAX = 32, 81, 100
NMS = dict(min_score=0.01, max_overlap=0.45, top_k=200)
WORKLOAD = ("SSD512 COCO-shaped synthetic: 24564 priors x 81 classes, <=100 GT/img, 32 img/GPU; step = "
            "MultiBoxLoss512 (SmoothL1 + CE hard-negative mining) fwd+bwd on a train batch + detect "
            "(0.01, 0.45, 200) on an eval batch")


class Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def make_cfg(device):
    return Cfg(device=device, n_classes=N_CLASSES, reg_weights=1.0, reg_loss="", cls_loss="",
               model={"box_type": "offset"}, focal_type="softmax")


# ---------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference, all host threads, bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(pri, train, evalb, n_train, n_eval):
    """One pass of the reference algorithm (oracle port) over n_train train images + n_eval eval images.
    Returns seconds per image for (loss fwd+bwd) and (detect)."""
    import torchvision
    from oracle import box_pipeline as O
    locs, scores, bx, lb = train
    l_c = locs[:n_train].clone().requires_grad_(True)
    s_c = scores[:n_train].clone().requires_grad_(True)
    t0 = time.perf_counter()
    loss = O.multibox_loss("s512", pri, l_c, s_c, bx[:n_train], lb[:n_train])
    loss.backward()
    t1 = time.perf_counter()
    O.detect(evalb[0][:n_eval].clone(), evalb[1][:n_eval], NMS["min_score"], NMS["max_overlap"], NMS["top_k"], pri,
             nms_fn=torchvision.ops.nms)
    t2 = time.perf_counter()
    return (t1 - t0) / n_train, (t2 - t1) / n_eval


def run_reference(args, rank, world):
    """--impl reference: rank 0 alone times the CPU path."""
    if rank != 0:
        return
    from shape_based_object_detection_b200 import priors as PR, synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pri = PR.PRIOR_TABLES[P_WORKLOAD]()
    n_train, n_eval = 4, 1
    train = synth.make_train_batch(pri, n_train, N_CLASSES, GMAX, 1234 + 2)
    evalb = synth.make_eval_batch(pri, n_eval, N_CLASSES, 4321)
    for _ in range(args.warmup):
        cpu_reference_step(pri, train, evalb, n_train, n_eval)
    t_tr, t_ev = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b = cpu_reference_step(pri, train, evalb, n_train, n_eval)
        t_tr.append(a)
        t_ev.append(b)
    wall = time.perf_counter() - t0
    t_tr.sort()
    t_ev.sort()
    per_img = t_tr[len(t_tr) // 2] + t_ev[len(t_ev) // 2]
    value = 1.0 / per_img
    sample = (f"{n_train} train images (loss fwd+bwd) + {n_eval} eval image (detect) per step of the same "
              f"workload; median over {args.steps} steps")
    print(json.dumps({
        "impl": "reference", "metric": "images/sec for assign+loss+NMS", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample,
                         "ms_per_image_loss_fwd_bwd": t_tr[len(t_tr) // 2] * 1e3,
                         "ms_per_image_detect": t_ev[len(t_ev) // 2] * 1e3},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------
def timed(fn, steps, sync):
    """CUDA-event time of `steps` calls of fn on the current stream, in ms per call."""
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / steps


def run_ours(args, rank, world, local_rank):
    import ctypes as C

    import torch.distributed as dist

    import shape_based_object_detection_b200 as S
    from shape_based_object_detection_b200 import _lib as L
    from shape_based_object_detection_b200 import priors as PR, synth
    from shape_based_object_detection_b200.models import MultiBoxLoss512
    from shape_based_object_detection_b200.models import utils as MU

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L.lib()
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    pri = PR.PRIOR_TABLES[P_WORKLOAD]()
    P = pri.size(0)
    N = N_PER_GPU
    # every rank owns its own 32 images (weak scaling); seeds differ per rank
    train = synth.make_train_batch(pri, N, N_CLASSES, GMAX, 1234 + 2 + 1000 * rank)
    evalb = synth.make_eval_batch(pri, N, N_CLASSES, 4321 + 1000 * rank)
    pin = lambda t: t.pin_memory()
    h_locs, h_scores = pin(train[0]), pin(train[1])
    h_bx, h_lb = [pin(b) for b in train[2]], [pin(l) for l in train[3]]
    h_elocs, h_escores = pin(evalb[0]), pin(evalb[1])
    pri_d = pri.to(dev)
    cfg = make_cfg(dev)
    crit = MultiBoxLoss512(pri_d, cfg)
    crit.process_group = group

    d_locs = h_locs.to(dev).requires_grad_(True)
    d_scores = h_scores.to(dev).requires_grad_(True)
    d_bx, d_lb = [b.to(dev) for b in h_bx], [l.to(dev) for l in h_lb]
    d_elocs, d_escores = h_elocs.to(dev), h_escores.to(dev)
    gt = S.pack_ground_truth(d_bx, d_lb, dev)

    # The train batch and the eval batch of a step are independent: by default the eval path runs on a
    # second CUDA stream (both captured in the same graph); --no-overlap serialises them.
    s_eval = torch.cuda.Stream(device=dev, priority=int(os.environ.get("SBOD_BENCH_SIDE_PRIO", "0")))

    def step_resident():
        d_locs.grad = None
        d_scores.grad = None
        cur = torch.cuda.current_stream()
        if args.overlap:
            s_eval.wait_stream(cur)
            with torch.cuda.stream(s_eval):
                out = S.detect_batched(d_elocs, d_escores, NMS["min_score"], NMS["max_overlap"], NMS["top_k"],
                                       pri_d)
        call = None
        if not args.overlap and args.presample:
            # the eval path's sampling pass (1/26 of the eval logits, one tile per CTA) is enqueued on
            # a side stream: it fills the SMs that the small-grid kernels of the train path leave idle
            call = S.detect_begin(d_elocs, d_escores, NMS["min_score"], NMS["max_overlap"], NMS["top_k"], pri_d,
                                  side_stream=s_eval)
        loss = crit.forward_packed(d_locs, d_scores, gt)  # GT packed once: inputs are resident
        loss.backward()
        if args.overlap:
            cur.wait_stream(s_eval)
        elif call is not None:
            out = S.detect_end(call)
        else:
            out = S.detect_batched(d_elocs, d_escores, NMS["min_score"], NMS["max_overlap"], NMS["top_k"], pri_d)
        return out

    def step_e2e():
        l = h_locs.to(dev, non_blocking=True).requires_grad_(True)
        s = h_scores.to(dev, non_blocking=True).requires_grad_(True)
        bx = [b.to(dev, non_blocking=True) for b in h_bx]
        lb = [x.to(dev, non_blocking=True) for x in h_lb]
        loss = crit(l, s, bx, lb)
        loss.backward()
        el = h_elocs.to(dev, non_blocking=True)
        es = h_escores.to(dev, non_blocking=True)
        out = S.detect_batched(el, es, NMS["min_score"], NMS["max_overlap"], NMS["top_k"], pri_d)
        res = (loss.detach().cpu(), out[0].cpu(), out[1].cpu(), out[2].cpu(), out[4].cpu())
        return res

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # The resident step is a fixed sequence of kernel launches: replay it as one CUDA graph so that the
    # Python / launch overhead (~0.5 ms per step, more than the kernels themselves) leaves the timed loop.
    graph = g = None
    eager_step = step_resident
    if args.graph:
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step_resident()
            g.replay()
            torch.cuda.synchronize()
            graph = g
            step_resident = g.replay
        except Exception as exc:  # capture is an optimisation, never a requirement
            print(f"bench.py: CUDA graph capture failed ({type(exc).__name__}: {exc}); running eagerly",
                  file=sys.stderr)
            step_resident = eager_step
            torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step_resident, args.steps, sync)
    clocks = sampler.stop() if rank == 0 else None

    # ---- stage timings + the dominant kernel alone (CUDA events on the launching stream) ----
    st = crit.last["state"]
    ms_fwd = timed(lambda: st.forward(), args.steps, sync)
    gl = torch.ones((), device=dev)
    g_locs_buf, g_scores_buf = torch.empty_like(st.locs), torch.empty_like(st.scores)
    ms_bwd = timed(lambda: st.backward_into(gl, g_locs_buf, g_scores_buf), args.steps, sync)
    ms_det = timed(lambda: S.detect_batched(d_elocs, d_escores, NMS["min_score"], NMS["max_overlap"],
                                            NMS["top_k"], pri_d), args.steps, sync)
    stage = L.lib().sbod_loss_forward_stage
    ms_match = timed(lambda: L.check(stage(C.byref(st.desc), 0, L.stream_ptr())), args.steps, sync)
    L.check(stage(C.byref(st.desc), 1, L.stream_ptr()))  # leave the workspace clean
    # the eval path's streaming kernel alone: build the descriptor once, then time stage 0 / stage 1 pairs
    det_desc = S.core.make_detect_desc(d_elocs, d_escores, NMS["min_score"], NMS["max_overlap"], NMS["top_k"], pri_d)
    dstage = L.lib().sbod_detect_stage
    ms_dscore = 0.0
    sync()
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.check(dstage(C.byref(det_desc["desc"]), 2, L.stream_ptr()))  # sampling pass (1/26 of the tiles)
        e0.record()
        L.check(dstage(C.byref(det_desc["desc"]), 3, L.stream_ptr()))  # the main score pass alone
        e1.record()
        L.check(dstage(C.byref(det_desc["desc"]), 1, L.stream_ptr()))
        torch.cuda.synchronize()
        ms_dscore += e0.elapsed_time(e1) / args.steps

    # ---- e2e: pinned host buffers in, results out, every step ----
    for _ in range(2):
        step_e2e()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed(step_e2e, e2e_steps, sync)

    t = torch.tensor([ms_step, ms_e2e, ms_fwd, ms_bwd, ms_det, ms_match, ms_dscore], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_e2e, ms_fwd, ms_bwd, ms_det, ms_match, ms_dscore = t.tolist()

    if rank == 0:
        T = int(gt[0].size(0))
        h2d = (h_locs.numel() + h_scores.numel() + h_elocs.numel() + h_escores.numel()) * 4 + T * 24
        d2h = 4 + N * NMS["top_k"] * (16 + 8 + 4) + N * 4
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # algorithmic bytes per launch of the two streaming kernels (DESIGN.md "Kernels"):
        #   train: logits once + priors once + GT, plus - when the backward follows - the zero-fill of
        #          the logits' gradient, which this kernel writes on the way (no separate fill kernel);
        #   eval:  logits once
        prefill = st.grad_scores is not None or bool(st.desc.grad_scores_prefill)
        alg_logits = N * P * N_CLASSES * 4 + P * 16 + T * 24
        alg_match = alg_logits + (N * P * N_CLASSES * 4 if prefill else 0)
        alg_dscore = N * P * N_CLASSES * 4
        per_kernel = {
            "match_lse_fast_kernel": {"ms": ms_match, "bytes": alg_match, "bytes_logits_only": alg_logits,
                                      "writes_gradient_zero_fill": prefill},
            "detect_score_fast_kernel": {"ms": ms_dscore, "bytes": alg_dscore},
        }
        for v in per_kernel.values():
            v["achieved_gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            v["frac"] = v["achieved_gbs"] / peak
        dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"])  # the dominant kernel of the step
        alg_bytes, ms_dom = per_kernel[dom]["bytes"], per_kernel[dom]["ms"]
        achieved = per_kernel[dom]["achieved_gbs"]
        traffic = None
        try:
            tj = json.load(open(os.path.join(REPO, "profiles", "roofline_traffic.json")))
            traffic = tj.get(dom + ("+zero_fill" if dom.startswith("match_lse") and prefill else ""), tj.get(dom))
        except Exception:
            pass
        # CPU baseline beside it (bounded sample, all host cores)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        c_train = (train[0][:2], train[1][:2], train[2][:2], train[3][:2])
        c_eval = (evalb[0][:1], evalb[1][:1])
        cpu_reference_step(pri, c_train, c_eval, 2, 1)
        a, b = cpu_reference_step(pri, c_train, c_eval, 2, 1)
        cpu_value = 1.0 / (a + b)
        line = {
            "metric": "images/sec for assign+loss+NMS", "value": N * world / (ms_step * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu": N, "l2": "inputs (2 x 255 MB logits) exceed the 126 MB L2",
                       "streams": "train and eval halves of the step on two CUDA streams" if args.overlap else (
                           "one stream + the eval sampling pass on a side stream" if args.presample else "one stream"),
                       "launch": "CUDA graph replay" if graph is not None else "eager Python launches",
                       "ms_loss_fwd": ms_fwd, "ms_loss_bwd": ms_bwd, "ms_detect": ms_det,
                       "images_per_s_loss_fwd": N * world / (ms_fwd * 1e-3),
                       "images_per_s_loss_fwd_bwd": N * world / ((ms_fwd + ms_bwd) * 1e-3),
                       "images_per_s_detect": N * world / (ms_det * 1e-3)},
            "e2e": {"value": N * world / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            # match_lse (+ gradient zero-fill), classify, mine, bwd_patch, detect sample, detect score,
            # detect nms, + the two (normally empty) fallback launches
            "gpu_launches": 9 * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650",
                         "ms_per_launch": ms_dom, "algorithmic_bytes_per_launch": alg_bytes,
                         "streaming_kernels": per_kernel},
            "cpu_baseline": {"value": cpu_value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "2 train images (loss fwd+bwd) + 1 eval image (detect), oracle port of the "
                                       "reference on the host cores, second of two passes",
                             "ms_per_image_loss_fwd_bwd": a * 1e3, "ms_per_image_detect": b * 1e3},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured step holds NCCL kernel nodes: release the graph before the communicator goes away
        step_resident = eager_step
        graph = g = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch the resident step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--no-presample", dest="presample", action="store_false",
                    help="keep the eval path's sampling pass on the main stream")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="run the train half and the eval half of a step one after the other on one stream "
                         "(default: on two CUDA streams inside the replayed graph - the two halves are independent, "
                         "and the small-grid kernels of one (mining, NMS: one CTA per image) fill the SMs the other "
                         "leaves idle: 0.253 ms vs 0.311 ms per step)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
