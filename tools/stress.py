#!/usr/bin/env python
"""Randomised shapes against the CPU oracle (GPU box only; slower and broader than the test suite): losses,
assignments and gradients of the three MultiBox variants + focal, and the eval path, on odd prior counts,
class counts on both sides of the compile-time specialisations, zero to several hundred objects per image.
usage: stress.py [n_cases] [seed]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import box_pipeline as O  # noqa: E402
from shape_based_object_detection_b200 import priors as PR, synth  # noqa: E402
from shape_based_object_detection_b200.models import MultiBoxLoss300, MultiBoxLoss512, RetinaFocalLoss  # noqa: E402
from shape_based_object_detection_b200.models import utils as MU  # noqa: E402


class Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def cfg(dev, n_classes, reg="", cls=""):
    return Cfg(device=dev, n_classes=n_classes, reg_weights=1.0, reg_loss=reg, cls_loss=cls,
               model={"box_type": "offset"}, focal_type="softmax")


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(seed)
    full = PR.ssd512_canonical_priors()
    variants = [(MultiBoxLoss512, "s512", "", ""), (MultiBoxLoss300, "s300", "", ""),
                (RetinaFocalLoss, "ret", "", "FOCAL"), (RetinaFocalLoss, "ret", "DIOU", "FOCAL")]
    bad = 0
    for case in range(n_cases):
        P = int(torch.randint(40, 9000, (1,), generator=gen))
        Cn = [2, 3, 4, 5, 21, 22, 81, 90, 130][int(torch.randint(0, 9, (1,), generator=gen))]
        N = int(torch.randint(1, 6, (1,), generator=gen))
        gmax = [1, 3, 20, 100, 300][int(torch.randint(0, 5, (1,), generator=gen))]
        Mod, variant, reg, cls = variants[case % len(variants)]
        idx = torch.randperm(full.size(0), generator=gen)[:P].sort().values
        pri = full[idx].contiguous()
        locs, scores, bx, lb = synth.make_train_batch(pri, N, Cn, gmax, 1000 + case)
        scale = [1.0, 0.1, 4.0][case % 3]
        scores = scores * scale
        l_c, s_c = locs.clone().requires_grad_(True), scores.clone().requires_grad_(True)
        want, parts = O.multibox_loss(variant, pri, l_c, s_c, bx, lb, reg_loss=reg, cls_loss=cls, want_parts=True)
        want.backward()
        crit = Mod(pri.to(dev), cfg(dev, Cn, reg, cls))
        l_d, s_d = locs.to(dev).requires_grad_(True), scores.to(dev).requires_grad_(True)
        loss = crit(l_d, s_d, [b.to(dev) for b in bx], [x.to(dev) for x in lb])
        loss.backward()
        st = crit.last["state"]
        ok = torch.equal(st.obj.cpu().long(), parts["obj"]) and torch.equal(st.ov.cpu(), parts["ov"])
        ok = ok and abs(loss.item() - want.item()) <= 1e-5 * abs(want.item())
        ok = ok and torch.allclose(l_d.grad.cpu(), l_c.grad, rtol=1e-4, atol=1e-7)
        g_ok = torch.allclose(s_d.grad.cpu(), s_c.grad, rtol=1e-4, atol=1e-7)
        if not g_ok:  # at most one near-tie swap of mined rows per image (see the mining tests)
            gm, wm = s_d.grad.cpu().abs().sum(2) > 0, s_c.grad.abs().sum(2) > 0
            both = gm & wm
            g_ok = int((gm ^ wm).sum()) <= 2 * N and torch.allclose(s_d.grad.cpu()[both], s_c.grad[both], rtol=1e-4, atol=1e-7)
        ok = ok and g_ok
        # eval path on the same priors
        elocs, escores = synth.make_eval_batch(pri, N, Cn, 2000 + case, bg_bias=[3.0, 6.0][case % 2])
        top_k = [5, 50, 200][case % 3]
        wantd = O.detect(elocs.clone(), escores, 0.01, 0.45, top_k, pri)
        gotd = MU.detect(elocs.to(dev), escores.to(dev), 0.01, 0.45, top_k, pri.to(dev), cfg(dev, Cn))
        d_ok = all(torch.equal(gotd[1][i].cpu(), wantd[1][i]) and
                   torch.allclose(gotd[2][i].cpu(), wantd[2][i], rtol=1e-5, atol=1e-8) and
                   torch.allclose(gotd[0][i].cpu(), wantd[0][i], rtol=1e-5, atol=1e-6) for i in range(N))
        print("case %2d %-5s %-5s P=%5d C=%3d N=%d G<=%3d scale %.1f: loss %s  detect %s" %
              (case, variant, reg or cls or "-", P, Cn, N, gmax, scale, "ok" if ok else "MISMATCH", "ok" if d_ok else "MISMATCH"),
              flush=True)
        bad += (not ok) + (not d_ok)
    print("stress: %d mismatches in %d cases" % (bad, n_cases))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
