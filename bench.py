#!/usr/bin/env python
"""bench.py — images/sec of the detection box pipeline (assign + loss fwd/bwd + NMS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]

--config picks one of BASELINE.json's five configs (default 2, the one the metric is quoted on):
  1 SSD300 VOC-shaped        8 732 priors x 21 classes, <= 20 GT, L1 + batch-global hard-negative mining
  2 SSD512 COCO-shaped      24 564 priors x 81 classes, <= 100 GT, SmoothL1 + per-image mining      (default)
  3 RetinaNet 640x640       76 725 anchors x 81 classes, softmax focal + GIoU, per-class top-1000 NMS
  4 RefineDet512 DETRAC     16 320 priors, ARM (2 classes) + ODM (4 classes), <= 200 dense GT, refined anchors
  5 FCOS 800x1333           22 300 locations x 81 columns, centre sampling + DIoU + NMS, 8 images per GPU
                            (= batch 64 sharded over 8 GPUs)
One step = the train path (targets + loss forward + backward) on a train batch AND the eval path (decode +
threshold + NMS + top-k) on an eval batch of the same per-GPU size. value = images/sec of the whole job with
inputs resident in HBM; e2e = the same through the public Python API with pinned HOST buffers (H2D of every
input and D2H of the results inside the timed region; the copies run on a copy stream into two slots of device
buffers, so the next step's copy overlaps this step's compute like a double-buffering data loader).
Multi-GPU: one process per GPU (torchrun), batch sharded by image, the only exchange is the loss sums
(four doubles per criterion; "weak" scaling); config 1's batch-global mining does not shard: replicas only.
The resident steps are replayed as CUDA graphs (--no-graph: eager launches) with the train halves and the eval
halves on two streams (--no-overlap: one stream). The two halves of a step work on independent batches, so the
streams are joined once per captured graph of U steps (U = 10 by default), not once per step: the eval half of
step i runs under the train half of step i+1 and the one-CTA-per-image tails of either half hide under the other
half's streaming kernels. The train halves of consecutive steps alternate between two streams as well (each with
its own criterion state, workspace and exchange lane - like two micro-batches in flight), so that the latency-bound
tail of one step's loss (classification, mining, sparse backward, the cross-GPU exchange of the sums) runs under the
next step's streaming kernel. That pays when the tail contains the exchange between GPUs (default with several
GPUs; on one GPU a single train stream is faster and is the default: --train-streams). The timed region still brackets EXACTLY K steps with a full synchronisation on both
sides; details.ms_per_step_joined is the same measurement with a join after every step (--join-every-step makes
it the headline).

--impl reference times the reference's own CPU implementation on the host cores (oracle/_ref when the build
placed it there, else the oracle port), on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "images/sec for assign+loss+NMS"

CONFIGS = {
    1: dict(name="ssd300_voc", priors="ssd300", C=21, gmax=20, n_per_gpu=32, nms=(0.01, 0.45, 200), shards=False,
            cpu_sample=(8, 2),
            workload="SSD300 VOC-shaped synthetic: 8732 priors x 21 classes, <=20 GT/img, 32 img/GPU; step = "
                     "MultiBoxLoss300 (L1 + CE batch-global hard-negative mining) fwd+bwd on a train batch + detect "
                     "(0.01, 0.45, 200) on an eval batch"),
    2: dict(name="ssd512_coco", priors="ssd512_canonical", C=81, gmax=100, n_per_gpu=32, nms=(0.01, 0.45, 200),
            shards=True, cpu_sample=(2, 1),
            workload="SSD512 COCO-shaped synthetic: 24564 priors x 81 classes, <=100 GT/img, 32 img/GPU; step = "
                     "MultiBoxLoss512 (SmoothL1 + CE hard-negative mining) fwd+bwd on a train batch + detect "
                     "(0.01, 0.45, 200) on an eval batch"),
    3: dict(name="retinanet640", priors="retinanet640", C=81, gmax=100, n_per_gpu=32, nms=(0.01, 0.45, 200),
            shards=True, cpu_sample=(1, 1),
            workload="RetinaNet 640x640 synthetic: 76725 anchors x 81 classes, <=100 GT/img, 32 img/GPU; step = "
                     "RetinaFocalLoss (softmax focal + GIoU) fwd+bwd on a train batch + detect (0.01, 0.45, 200) "
                     "with the per-class top-1000 candidate cap on an eval batch"),
    4: dict(name="refinedet512_detrac", priors="refinedet512", C=4, gmax=200, n_per_gpu=32, nms=(0.01, 0.45, 200),
            shards=True, cpu_sample=(2, 1),
            workload="RefineDet512 DETRAC-shaped synthetic: 16320 priors, ARM 2 classes + ODM 4 classes, <=200 dense "
                     "GT/img, 32 img/GPU; step = RefineDetLoss (ARM + ODM against refined anchors) fwd+bwd on a train "
                     "batch + offset2bbox + detect_refine (0.01, 0.45, 200, second NMS 0.7) on an eval batch"),
    5: dict(name="fcos_800x1333", priors=None, C=81, gmax=100, n_per_gpu=8, nms=(0.05, 0.45, 100), shards=True,
            cpu_sample=(1, 1),
            workload="FCOS 800x1333 COCO-shaped synthetic: 22300 locations x (80 classes + centerness), <=100 GT/img, "
                     "8 img/GPU (batch 64 over 8 GPUs); step = FCOSLoss (centre-sampling assignment + sigmoid focal + "
                     "centerness-weighted DIoU + BCE) fwd+bwd on a train batch + postprocess + detect (0.05, 0.45, 100) "
                     "on an eval batch; reference FCOS code does not run: parity unpinned, CPU arm = oracle port"),
}


def config_entry(cid):
    """The `config` object of the JSON line: identical on both arms (ours / reference) by construction."""
    cf = CONFIGS[cid]
    return {"workload": cf["workload"], "config_id": cid, "images_per_gpu": cf["n_per_gpu"], "l2": L2_NOTE[cid]}


# per-step input bytes of a config vs the 126 MB L2 (timing rule: say whether the inputs exceed it)
L2_NOTE = {
    1: "inputs of a step (2 x 23.5 MB logits) FIT the 126 MB L2: this config is latency / issue bound, not HBM bound",
    2: "inputs of a step (2 x 255 MB logits) exceed the 126 MB L2",
    3: "inputs of a step (2 x 796 MB logits) exceed the 126 MB L2",
    4: "inputs of a step (< 40 MB) FIT the 126 MB L2: this config is FP32-issue / latency bound (P x G pair evaluations)",
    5: "inputs of a step (2 x 58 MB logits) are of the order of the 126 MB L2; consecutive steps alternate train and "
       "eval buffers (231 MB touched per step)",
}


class Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def make_cfg(device, n_classes, reg="", cls=""):
    return Cfg(device=device, n_classes=n_classes, reg_weights=1.0, reg_loss=reg, cls_loss=cls,
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
# synthetic inputs of a config (CPU tensors; SURVEY §8d distributions)
# ---------------------------------------------------------------------------------------------
def make_inputs(cid, n, seed_train, seed_eval):
    """Returns (priors_or_locations, train dict, eval dict) of CPU tensors."""
    from shape_based_object_detection_b200 import priors as PR, synth
    cf = CONFIGS[cid]
    Cn = cf["C"]
    if cid == 5:
        from shape_based_object_detection_b200.models import compute_location
        locations = compute_location(image_size=(800, 1333))
        P = sum(l.size(0) for l in locations)
        gen = torch.Generator().manual_seed(seed_train)
        bx, lb = synth.make_gt(n, cf["gmax"], Cn, gen)
        train = dict(tensors=[torch.rand((n, P, 4), generator=gen) * 0.3 + 0.01, torch.randn((n, P, Cn), generator=gen),
                              torch.randn((n, P), generator=gen)], boxes=bx, labels=lb)
        gen = torch.Generator().manual_seed(seed_eval)
        ev = dict(tensors=[torch.rand((n, P, 4), generator=gen) * 0.3 + 0.01,
                           torch.randn((n, P, Cn), generator=gen) * 2.0 - 3.0, torch.randn((n, P), generator=gen)])
        return locations, train, ev
    pri = PR.PRIOR_TABLES[cf["priors"]]()
    P = pri.size(0)
    if cid == 4:
        gen = torch.Generator().manual_seed(seed_train)
        bx, lb = synth.make_gt(n, cf["gmax"], Cn, gen, dense=True)
        train = dict(tensors=[torch.randn((n, P, 4), generator=gen) * 0.1, torch.randn((n, P, 2), generator=gen) * 2.0,
                              torch.randn((n, P, 4), generator=gen) * 0.1, torch.randn((n, P, Cn), generator=gen)],
                     boxes=bx, labels=lb)
        gen = torch.Generator().manual_seed(seed_eval)
        sc = torch.randn((n, P, Cn), generator=gen) * 2.0
        sc[:, :, 0] += 8.0  # as synth.make_eval_batch: a few percent of the (prior, class) scores pass min_score
        ev = dict(tensors=[torch.randn((n, P, 4), generator=gen) * 0.3, torch.randn((n, P, 4), generator=gen) * 0.3, sc],
                  arm_scores=torch.randn((n, P, 2), generator=gen) * 2.0)
        return pri, train, ev
    locs, scores, bx, lb = synth.make_train_batch(pri, n, Cn, cf["gmax"], seed_train)
    elocs, escores = synth.make_eval_batch(pri, n, Cn, seed_eval)
    return pri, dict(tensors=[locs, scores], boxes=bx, labels=lb), dict(tensors=[elocs, escores])


def algorithmic_bytes(cid, n, P, T):
    """SURVEY §8(d) per-step algorithmic bytes of a config (train fwd+bwd + eval), for the step-level check."""
    Cn = CONFIGS[cid]["C"]
    if cid == 4:  # ARM (2 classes) + ODM (Cn classes): logits, locs, per-image anchors, grads
        per = P * ((2 + Cn) * 4 * 3 + 16 * 4 + 16)
        return n * per + P * 16 + T * 24 * 2 + n * P * (Cn * 4 + 32)
    if cid == 5:
        return n * P * ((Cn * 4 + 16 + 4) * 3) + T * 24 + n * P * (Cn * 4 * 3 + 20)
    train = n * P * (Cn * 4 * 3 + 16 * 2) + P * 16 + T * 24
    ev = n * P * (Cn * 4 + 16) + P * 16
    return train + ev


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref) or the oracle port, bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_arm(cid, reps, warm):
    """Times the CPU implementation of config `cid` on a bounded sample. Returns the cpu_baseline dict."""
    from oracle import ref_runner as RR
    cf = CONFIGS[cid]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_train, n_eval = cf["cpu_sample"]
    pri, train, ev = make_inputs(cid, max(n_train, n_eval), 1234 + cid, 4321 + cid)
    runner = RR.Runner(cid, pri, cf["C"])
    tt = [t[:n_train] for t in train["tensors"]]
    et = [t[:n_eval] for t in ev["tensors"]]
    keep = (ev["arm_scores"][:n_eval, :, 1] > 0.01) if cid == 4 else None

    def once():
        t0 = time.perf_counter()
        runner.train(tt, train["boxes"][:n_train], train["labels"][:n_train])
        t1 = time.perf_counter()
        if cid == 4:
            runner.eval(et, cf["nms"], keep)
        else:
            runner.eval(et, cf["nms"])
        t2 = time.perf_counter()
        return (t1 - t0) / n_train, (t2 - t1) / n_eval

    for _ in range(warm):
        once()
    a, b = [], []
    t0 = time.perf_counter()
    for _ in range(reps):
        x, y = once()
        a.append(x)
        b.append(y)
    wall = time.perf_counter() - t0
    a.sort()
    b.sort()
    ta, tb = a[len(a) // 2], b[len(b) // 2]
    value = 1.0 / (ta + tb)
    src = ("the reference's own code (oracle/_ref, unmodified)" if runner.kind == "reference"
           else "the oracle port of the reference (oracle/box_pipeline.py)")
    sample = (f"{n_train} train image(s) (loss fwd+bwd) + {n_eval} eval image(s) (detect) of the same workload per "
              f"pass, {src}, median of {reps} pass(es) after {warm} warm-up")
    return {"value": value, "unit": "images/s", "cores": cores, "kind": runner.kind, "sample": sample,
            "ms_per_image_loss_fwd_bwd": ta * 1e3, "ms_per_image_detect": tb * 1e3}, wall


def run_reference(args, rank, world):
    """--impl reference: rank 0 alone times the CPU path."""
    if rank != 0:
        return
    cf = CONFIGS[args.config]
    base, wall = cpu_arm(args.config, max(args.steps, 1), max(args.warmup, 1))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_entry(args.config),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------
def timed(fn, steps, sync):
    """CUDA-event time of `steps` calls of fn on the current stream, in ms per call (after one untimed call:
    scratch buffers are per stream and are allocated on first use)."""
    fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / steps


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPU cores next to its GPU before any pinned host buffer is allocated, so that the
    H2D copies of the e2e leg do not cross sockets (eight ranks copying at once otherwise contend on the host)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class Workload:
    """Device-side state of one config: criteria, resident tensors, the train / eval halves of a step."""

    def __init__(self, cid, dev, rank, group):
        import shape_based_object_detection_b200 as S
        from shape_based_object_detection_b200 import models as M
        self.S, self.cid, self.dev, self.cf = S, cid, dev, CONFIGS[cid]
        cf = self.cf
        self.N = cf["n_per_gpu"]
        pri, train, ev = make_inputs(cid, self.N, 1234 + cid + 1000 * rank, 4321 + cid + 1000 * rank)
        self.h_train = [t.pin_memory() for t in train["tensors"]]
        self.h_eval = [t.pin_memory() for t in ev["tensors"]]
        self.h_boxes = [b.pin_memory() for b in train["boxes"]]
        self.h_labels = [l.pin_memory() for l in train["labels"]]
        self.T = sum(int(b.size(0)) for b in train["boxes"])
        self.d_train = [t.to(dev).requires_grad_(True) for t in self.h_train]
        self.d_eval = [t.to(dev) for t in self.h_eval]
        d_boxes, d_labels = [b.to(dev) for b in self.h_boxes], [l.to(dev) for l in self.h_labels]
        self.gt = S.pack_ground_truth(d_boxes, d_labels, dev)
        from shape_based_object_detection_b200.dataset.collate import PackedGT
        self.d_packed = PackedGT.from_lists(train["boxes"], train["labels"]).to(dev)  # resident CSR ground truth
        grp = group if cf["shards"] else None
        if cid == 5:
            self.locations = [l.to(dev) for l in pri]
            self.P = sum(l.size(0) for l in pri)
            self.crit = M.FCOSLoss(self.locations, make_cfg(dev, cf["C"]), image_size=(800, 1333))
        else:
            self.pri = pri.to(dev)
            self.P = pri.size(0)
            if cid == 1:
                self.crit = M.MultiBoxLoss300(self.pri, make_cfg(dev, cf["C"]))
            elif cid == 2:
                self.crit = M.MultiBoxLoss512(self.pri, make_cfg(dev, cf["C"]))
            elif cid == 3:
                self.crit = M.RetinaFocalLoss(self.pri, make_cfg(dev, cf["C"], "GIOU", "FOCAL"))
                self.crit.extended_reg_losses = True  # GIoU is an explicit opt-in (the reference only knows DIoU)
            else:
                self.crit = M.RefineDetLoss(self.pri, make_cfg(dev, cf["C"]))
                self.h_arm_scores = ev["arm_scores"].pin_memory()
                self.d_keep = (self.h_arm_scores.to(dev)[:, :, 1] > 0.01)  # RefineDet512.py:639 (raw logit > theta)
        self.crit.process_group = grp
        self.d_boxes, self.d_labels = d_boxes, d_labels
        # a second instance of the criterion for the micro-batch in flight on the second train stream: its own
        # state and its own exchange lane (the loss sums of concurrent criteria must not share mailboxes)
        import copy
        self.crit2 = copy.copy(self.crit)
        for name in ("last", "last_arm", "last_odm"):
            if hasattr(self.crit2, name):
                setattr(self.crit2, name, {})
        self.crit2.exchange_lane = 1

    # ---- halves of a step on resident inputs ----
    def train_half(self, tensors=None, boxes=None, labels=None, slot=0):
        ts = tensors if tensors is not None else self.d_train
        crit = self.crit2 if slot else self.crit
        for t in ts:
            t.grad = None
        if self.cid in (1, 2, 3) and tensors is None:
            loss = crit.forward_packed(ts[0], ts[1], self.gt)  # GT packed once: inputs are resident
        elif tensors is None:
            loss = crit(*ts, self.d_packed, None)  # (RefineDet / FCOS take the packed batch as `boxes`)
        else:
            loss = crit(*ts, boxes if boxes is not None else self.d_boxes,
                        labels if labels is not None else self.d_labels)
        loss.backward()
        return loss

    def forward_only(self):
        """The criterion on the resident train batch without autograd (no backward follows)."""
        with torch.no_grad():
            ts = [t.detach() for t in self.d_train]
            if self.cid in (1, 2, 3):
                return self.crit.forward_packed(ts[0], ts[1], self.gt)
            return self.crit(*ts, self.d_packed, None)

    def eval_half(self, tensors=None, keep=None):
        S, cf = self.S, self.cf
        ts = tensors if tensors is not None else self.d_eval
        ms, mo, tk = cf["nms"]
        if self.cid in (1, 2):
            return S.detect_batched(ts[0], ts[1], ms, mo, tk, self.pri)
        if self.cid == 3:
            return S.detect_batched(ts[0], ts[1], ms, mo, tk, self.pri, pre_nms_topk=1000)
        if self.cid == 4:
            from shape_based_object_detection_b200.models import offset2bbox
            boxes = offset2bbox(ts[0], ts[1], self.pri)
            return S.detect_batched(boxes, ts[2], ms, mo, tk, self.pri, box_type="corner", clamp_inplace=True,
                                    prior_keep=keep if keep is not None else self.d_keep, second_nms_thr=0.7)
        from shape_based_object_detection_b200.models import fcos_postprocess
        bl, sc = fcos_postprocess(ts[0], ts[1], ts[2], self.locations)
        return S.detect_batched(bl, sc, ms, mo, tk, None, act="none", box_type="corner", clamp_inplace=True)

    def h2d_bytes(self):
        n = sum(t.numel() * t.element_size() for t in self.h_train + self.h_eval) + self.T * 24
        if self.cid == 4:
            n += self.h_arm_scores.numel() * 4
        return n

    def d2h_bytes(self):
        return 4 + self.N * self.cf["nms"][2] * (16 + 8 + 4) + self.N * 4

    def launches_per_step(self, world):
        """Kernels of libsbod.so launched per resident step (counted from the call sequence of each path)."""
        ex = 1 if (world > 1 and self.cf["shards"]) else 0  # finalize after the all-reduce
        det = 3  # bound pass, refine, NMS
        return {1: 4 + det,            # match_lse, classify, mine, bwd_rows
                2: 4 + ex + det,       # match_lse, classify, mine, bwd_rows
                3: 4 + ex + det,       # match_lse, classify, mine, loss_bwd (dense)
                4: 2 * (4 + ex) + 2 + det + 1,  # ARM + ODM, decode_arm + easy-negative mask, offset2bbox
                5: 4 + ex + det + 1}[self.cid]  # assign, terms, finalize, backward terms, postprocess


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    import shape_based_object_detection_b200 as S
    from shape_based_object_detection_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    numa_cpus = bind_to_gpu_numa(local_rank)
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

    cid = args.config
    cf = CONFIGS[cid]
    W = Workload(cid, dev, rank, group)
    N, P = W.N, W.P
    s_eval = torch.cuda.Stream(device=dev)
    s_train2 = torch.cuda.Stream(device=dev)
    s_copy = torch.cuda.Stream(device=dev)

    def steps_resident(u=1):
        """u steps: the train halves on the current stream, the eval halves on the eval stream, one join."""
        cur = torch.cuda.current_stream()
        out = None
        if args.overlap:
            two = args.train_streams > 1 and u > 1
            s_eval.wait_stream(cur)
            if two:
                s_train2.wait_stream(cur)
            with torch.cuda.stream(s_eval):
                for _ in range(u):
                    out = W.eval_half()
            for k in range(u):
                if two and k % 2 == 1:  # odd steps' train halves on the second train stream (own criterion state)
                    with torch.cuda.stream(s_train2):
                        W.train_half(slot=1)
                else:
                    W.train_half()
            cur.wait_stream(s_eval)
            if two:
                cur.wait_stream(s_train2)
        else:
            for _ in range(u):
                W.train_half()
                out = W.eval_half()
        return out

    def step_resident():
        return steps_resident(1)

    # steps per captured graph (one stream join per graph); K must be a whole number of graphs
    unit = 1
    if not args.join_every_step:
        unit = next(u for u in (10, 8, 6, 5, 4, 3, 2, 1) if args.steps % u == 0)

    # ---- e2e: pinned host buffers in, results out, through the public API ----
    # Two slots of device input buffers (a data loader's double buffering): the H2D copy of step i+1 is enqueued on
    # the copy stream before the host blocks on the results of step i, so it runs under step i's compute and D2H.
    # Every step's inputs are copied from pinned host memory and every step's results are read back to the host
    # inside the timed loop. The ground truth travels packed (dataset.collate.PackedGT: one copy per batch).
    from shape_based_object_detection_b200.dataset.collate import PackedGT
    h_packed = PackedGT.from_lists(W.h_boxes, W.h_labels)  # pinned

    class Slot:
        def __init__(self):
            self.train = [torch.empty_like(t, device=dev).requires_grad_(True) for t in W.h_train]
            self.eval = [torch.empty_like(t, device=dev) for t in W.h_eval]
            self.gt = PackedGT(torch.empty_like(h_packed.buf, device=dev), h_packed.n_images, h_packed.total, h_packed.gmax)
            self.arm = torch.empty_like(W.h_arm_scores, device=dev) if cid == 4 else None
            self.ev_train, self.ev_eval, self.done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

    slots = [Slot(), Slot()]
    h_res = None  # pinned result buffers, allocated on first use

    def e2e_copy_in(slot):
        with torch.cuda.stream(s_copy):
            s_copy.wait_event(slot.done)  # the step that used this slot last has finished with it
            with torch.no_grad():
                for d, h in zip(slot.train, W.h_train):
                    d.copy_(h, non_blocking=True)
                slot.gt.buf.copy_(h_packed.buf, non_blocking=True)
                slot.ev_train.record(s_copy)
                for d, h in zip(slot.eval, W.h_eval):
                    d.copy_(h, non_blocking=True)
                if slot.arm is not None:
                    slot.arm.copy_(W.h_arm_scores, non_blocking=True)
                slot.ev_eval.record(s_copy)

    def e2e_compute(slot):
        nonlocal h_res
        cur = torch.cuda.current_stream()
        cur.wait_event(slot.ev_train)
        loss = W.train_half(slot.train, slot.gt, None)
        cur.wait_event(slot.ev_eval)
        keep = (slot.arm[:, :, 1] > 0.01) if slot.arm is not None else None
        out = W.eval_half(slot.eval, keep)
        outs = (loss.detach().reshape(1), out[0], out[1], out[2], out[4])
        if h_res is None:
            h_res = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
        for h, o in zip(h_res, outs):
            h.copy_(o, non_blocking=True)
        slot.done.record(cur)
        slot.done.synchronize()  # the host has the step's loss and detections
        return h_res

    def run_e2e(steps):
        e2e_copy_in(slots[0])
        res = None
        for i in range(steps):
            if i + 1 < steps:
                e2e_copy_in(slots[(i + 1) & 1])
            res = e2e_compute(slots[i & 1])
        return res

    for _ in range(max(args.warmup, 3)):
        step_resident()
    steps_resident(unit)  # (every stream / criterion of the captured graph has run once: workspaces, communicators)
    torch.cuda.synchronize()
    graph = g = g1 = None
    eager_step = step_resident
    replay = lambda: steps_resident(unit)  # noqa: E731
    replay1 = step_resident
    if args.graph:
        try:
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                step_resident()
            g1.replay()
            g = g1
            if unit > 1:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    steps_resident(unit)
                g.replay()
            torch.cuda.synchronize()
            graph = g
            replay, replay1 = g.replay, g1.replay
        except Exception as exc:  # capture is an optimisation, never a requirement
            print(f"bench.py: CUDA graph capture failed ({type(exc).__name__}: {exc}); running eagerly",
                  file=sys.stderr)
            torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(replay, args.steps // unit, sync) / unit
    clocks = sampler.stop() if rank == 0 else None
    ms_step_joined = timed(replay1, args.steps, sync) if unit > 1 else ms_step

    # ---- halves and single kernels alone (CUDA events on the launching stream) ----
    def graphed(fn):
        """fn captured in a CUDA graph of its own (device time without Python launch gaps); eager if that fails"""
        if not args.graph:
            return fn, None
        try:
            fn()
            torch.cuda.synchronize()
            gg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gg):
                fn()
            gg.replay()
            torch.cuda.synchronize()
            return gg.replay, gg
        except Exception:
            torch.cuda.synchronize()
            return fn, None

    def repeated_ms(launch, cleanup, reps=10):
        """Average duration of ONE kernel: `reps` launches back to back (a graph of them when graphs are on), so
        that the launch gap of a lone launch between two events does not count as kernel time. cleanup() restores
        the workspace contract afterwards (the later stages consume what the kernel accumulated)."""
        def many():
            for _ in range(reps):
                launch()
        run, gg = graphed(many)
        ms = timed(run, max(args.steps // reps, 3), sync) / reps
        cleanup()
        sync()
        return ms

    run_train, g_train = graphed(lambda: W.train_half())
    run_eval, g_eval = graphed(lambda: W.eval_half())
    ms_train = timed(run_train, args.steps, sync)
    ms_eval = timed(run_eval, args.steps, sync)
    kernels = {}  # name -> (ms per launch, algorithmic bytes per launch, note)
    Cn = cf["C"]
    dstage = L.lib().sbod_detect_stage
    if cid in (1, 2, 3, 4):
        st = W.crit.last["state"] if cid != 4 else W.crit.last_odm["state"]
        Cs = st.C
        stage = L.lib().sbod_loss_forward_stage
        prefill = bool(st.desc.grad_scores_prefill)
        ms_match = ms_cm = 0.0  # pairs: classify_mine consumes what the match kernel left in the workspace
        sync()
        for _ in range(args.steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            L.check(stage(C.byref(st.desc), 0, L.stream_ptr()))
            ev[1].record()
            L.check(stage(C.byref(st.desc), 1, L.stream_ptr()))
            ev[2].record()
            torch.cuda.synchronize()
            ms_match += ev[0].elapsed_time(ev[1]) / args.steps
            ms_cm += ev[1].elapsed_time(ev[2]) / args.steps
        alg = N * P * Cs * 4 + P * 16 + W.T * 24 + (N * P * 16 if cid == 4 else 0)
        ms_match_pair = ms_match
        ms_match = repeated_ms(lambda: L.check(stage(C.byref(st.desc), 0, L.stream_ptr())),
                               lambda: L.check(stage(C.byref(st.desc), 1, L.stream_ptr())))
        kernels["match_lse_fast_kernel"] = dict(ms=ms_match, bytes=alg + (N * P * Cs * 4 if prefill else 0),
                                                bytes_logits_only=alg, writes_gradient_zero_fill=prefill,
                                                ms_single_launch_between_events=ms_match_pair)
        kernels["classify_kernel+mine_kernel"] = dict(ms=ms_cm, bytes=N * P * 17)
        run_fwd, g_fwd = graphed(lambda: st.forward())
        ms_fwd = timed(run_fwd, args.steps, sync)
        # the forward alone as an inference of the loss would run it: no backward follows, so the streaming kernel
        # does not zero-fill a gradient buffer
        ms_fwd_only = None
        if prefill:
            W.forward_only()
            st0 = W.crit.last["state"] if cid != 4 else W.crit.last_odm["state"]
            if not bool(st0.desc.grad_scores_prefill):
                run_fwd0, g_fwd0 = graphed(lambda: st0.forward())
                ms_fwd_only = timed(run_fwd0, args.steps, sync)
        gl = torch.ones((), device=dev)
        g_l, g_s = torch.empty_like(st.locs), torch.empty_like(st.scores)
        ms_bwd = timed(lambda: st.backward_into(gl, g_l, g_s), args.steps, sync)
        if cid == 3:  # focal: dense backward, one kernel (logits in, gradient out)
            kernels["loss_bwd_kernel"] = dict(ms=ms_bwd, bytes=2 * N * P * Cs * 4 + N * P * 16)
        if cid == 4:
            pairs = P * W.T  # prior x object pairs of one assignment pass
            kernels["match_lse_fast_kernel"]["pair_evals_per_launch"] = pairs
            kernels["match_lse_fast_kernel"]["pair_evals_per_s"] = pairs / (ms_match * 1e-3)
    else:
        ms_fwd = ms_bwd = ms_fwd_only = None
    # eval path: the streaming bound pass alone (stage 2), then refine (3) and NMS (1)
    if True:
        if cid in (1, 2, 3):
            det = S.core.make_detect_desc(W.d_eval[0], W.d_eval[1], *cf["nms"], W.pri,
                                          pre_nms_topk=1000 if cid == 3 else 0)
            det_C = Cn
        elif cid == 4:
            from shape_based_object_detection_b200.models import offset2bbox
            boxes4 = offset2bbox(W.d_eval[0], W.d_eval[1], W.pri)
            det = S.core.make_detect_desc(boxes4, W.d_eval[2], *cf["nms"], W.pri, box_type="corner", clamp_inplace=True,
                                          prior_keep=W.d_keep, second_nms_thr=0.7)
            det_C = Cn
        else:
            from shape_based_object_detection_b200.models import fcos_postprocess
            bl5, sc5 = fcos_postprocess(W.d_eval[0], W.d_eval[1], W.d_eval[2], W.locations)
            det = S.core.make_detect_desc(bl5, sc5, *cf["nms"], None, act="none", box_type="corner", clamp_inplace=True)
            det_C = Cn
        acc = [0.0, 0.0, 0.0]
        sync()
        for _ in range(args.steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            L.check(dstage(C.byref(det["desc"]), 2, L.stream_ptr()))
            ev[1].record()
            L.check(dstage(C.byref(det["desc"]), 3, L.stream_ptr()))
            ev[2].record()
            L.check(dstage(C.byref(det["desc"]), 1, L.stream_ptr()))
            ev[3].record()
            torch.cuda.synchronize()
            for i in range(3):
                acc[i] += ev[i].elapsed_time(ev[i + 1]) / args.steps
        ms_bound = repeated_ms(lambda: L.check(dstage(C.byref(det["desc"]), 2, L.stream_ptr())),
                               lambda: L.check(dstage(C.byref(det["desc"]), 4, L.stream_ptr())))
        kernels["detect_bound_kernel"] = dict(ms=ms_bound, bytes=N * P * det_C * 4 + N * P * 4,
                                              ms_single_launch_between_events=acc[0])
        kernels["detect_refine_kernel"] = dict(ms=acc[1], bytes=None)
        kernels["detect_nms_kernel"] = dict(ms=acc[2], bytes=None)
    if cid == 5:
        from shape_based_object_detection_b200.models import fcos_postprocess
        ms_pp = timed(lambda: fcos_postprocess(W.d_eval[0], W.d_eval[1], W.d_eval[2], W.locations), args.steps, sync)
        kernels["fcos_postprocess_kernel"] = dict(ms=ms_pp, bytes=N * P * (Cn * 4 * 2 + 16 * 2 + 4))

    # ---- e2e ----
    run_e2e(3)
    e2e_steps = max(4, min(args.steps, 10))
    sync()
    e2e_t0, e2e_t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_t0.record()
    run_e2e(e2e_steps)
    e2e_t1.record()
    sync()
    ms_e2e = e2e_t0.elapsed_time(e2e_t1) / e2e_steps
    # H2D alone (per rank, all ranks copying at once): is the e2e number a host-side limit?
    def copy_only():
        return [t.to(dev, non_blocking=True) for t in W.h_train + W.h_eval]
    ms_h2d = timed(copy_only, e2e_steps, sync)
    h2d_gbs = sum(t.numel() * t.element_size() for t in W.h_train + W.h_eval) / (ms_h2d * 1e-3) / 1e9

    vals = [ms_step, ms_e2e, ms_train, ms_eval, ms_h2d, ms_step_joined] + [k["ms"] for k in kernels.values()]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    vals = t.tolist()
    ms_step, ms_e2e, ms_train, ms_eval, ms_h2d_max, ms_step_joined = vals[:6]
    for k, v in zip(kernels.values(), vals[6:]):
        k["ms"] = v
    h2d_all = [None] * world
    if world > 1:
        dist.all_gather_object(h2d_all, round(h2d_gbs, 2))
    else:
        h2d_all = [round(h2d_gbs, 2)]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        for v in kernels.values():
            if v.get("bytes"):
                v["achieved_gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
                v["frac"] = v["achieved_gbs"] / peak
        with_bytes = {k: v for k, v in kernels.items() if v.get("bytes")}
        dom = max(with_bytes, key=lambda k: with_bytes[k]["ms"])  # the dominant kernel of the step
        traffic = None
        try:
            tj = json.load(open(os.path.join(REPO, "profiles", "roofline_traffic.json")))
            traffic = tj.get("config%d" % cid, {}).get(dom)
        except Exception:
            pass
        cpu, _ = cpu_arm(cid, 1, 1)
        step_bytes = algorithmic_bytes(cid, N, P, W.T)
        line = {
            "metric": METRIC, "value": N * world / (ms_step * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_entry(cid),
            "details": {"streams": ("train halves on %d stream(s) (consecutive steps alternate), eval halves on one more, "
                                    "joined once per %d step(s)" % (args.train_streams if unit > 1 else 1, unit))
                        if args.overlap else "one stream",
                        "launch": ("CUDA graph replay, %d step(s) per graph" % unit) if graph is not None
                        else "eager Python launches",
                        "ms_per_step_joined": ms_step_joined,
                        "sharding": "batch sharded by image, one all-reduce of the loss sums per criterion" if cf["shards"]
                        else "independent replicas (batch-global mining does not shard)",
                        "ms_train_half": ms_train, "ms_eval_half": ms_eval, "ms_loss_fwd": ms_fwd,
                        "ms_loss_fwd_without_gradient_zero_fill": ms_fwd_only,
                        "ms_loss_bwd_incl_zero_fill": ms_bwd, "ms_detect": ms_eval,
                        "images_per_s_train_half": N * world / (ms_train * 1e-3),
                        "images_per_s_detect": N * world / (ms_eval * 1e-3),
                        "step_algorithmic_bytes": step_bytes, "step_gbs": step_bytes / (ms_step * 1e-3) / 1e9,
                        "numa_cpus_bound": numa_cpus},
            "e2e": {"value": N * world / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": W.h2d_bytes(),
                    "d2h_bytes_per_step": W.d2h_bytes(), "ms_per_step": ms_e2e,
                    "copy": "H2D on a copy stream into two slots of device buffers (the next step's copy runs under this step's compute and D2H); ground truth packed, results into pinned host buffers",
                    "h2d_gbs_per_rank_all_ranks_copying": h2d_all},
            "gpu_launches": W.launches_per_step(world) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": kernels[dom]["achieved_gbs"] / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650",
                         "ms_per_launch": kernels[dom]["ms"], "algorithmic_bytes_per_launch": kernels[dom]["bytes"],
                         "kernels": kernels},
            "cpu_baseline": cpu,
        }
        if cid in (1, 4):
            line["roofline"]["note"] = ("the working set of this config fits the L2: its kernels are latency / issue "
                                        "bound, the HBM fraction is reported for completeness")
        print(json.dumps(line), flush=True)
    if world > 1:
        graph = g = g1 = g_train = g_eval = run_train = run_eval = None  # the captured steps hold NCCL kernel nodes: release it before the communicator goes away
        replay = eager_step
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 300; 10 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config (1 SSD300, 2 SSD512 [default], 3 RetinaNet-640, 4 RefineDet512, 5 FCOS)")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch the resident step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--train-streams", type=int, default=None, choices=[1, 2],
                    help="train halves of consecutive steps alternate between this many streams (2: the cross-GPU "
                         "exchange of one step's loss sums waits under the next step's streaming kernel; default: 1 on "
                         "one GPU - measured 0.193 vs 0.202 ms per step - and 2 on several - 0.201 vs 0.210 ms at two)")
    ap.add_argument("--join-every-step", action="store_true",
                    help="join the train and eval streams after every step instead of once per captured graph of steps")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="run the train half and the eval half of a step one after the other on one stream")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.steps is None:
        args.steps = 10 if args.impl == "reference" else 300
    if args.train_streams is None:
        args.train_streams = 2 if world > 1 else 1
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
