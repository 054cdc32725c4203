"""Multi-GPU hardware check (run under torchrun with >= 2 ranks, one GPU each; driven by
tests/test_gpu_multi.py and by hand: `torchrun --nproc-per-node 2 tests/multi_gpu_check.py`).

The only exchange of the path is the loss sums of a batch sharded by image. Two
implementations: (A) the in-kernel NVLink mailbox exchange (csrc/comm.cuh, default) and (B) an NCCL
all-reduce between the forward and sbod_loss_finalize (SBOD_PEER_EXCHANGE=0) — the checked reference of the
exchange. Checks, per rank: A == B bit for bit; both equal the single-process run of the WHOLE batch (loss
1e-6, this rank's gradient rows 1e-6); the stand-alone mailbox all-reduce equals dist.all_reduce over many
epochs; the sharded step replays correctly from a CUDA graph; FCOS (six sums) likewise.
"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


class Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    from shape_based_object_detection_b200 import parallel, priors as PR, synth
    from shape_based_object_detection_b200.models import FCOSLoss, MultiBoxLoss512, RefineDetLoss, compute_location
    from shape_based_object_detection_b200.parallel import shard_range

    def cfg(n_classes):
        return Cfg(device=dev, n_classes=n_classes, reg_weights=1.0, reg_loss="", cls_loss="", model={"box_type": "offset"},
                   focal_type="softmax")

    # ---- stand-alone mailbox all-reduce vs NCCL over many epochs ----
    px = parallel.peer_exchange(group, dev)
    assert px is not None, "peer exchange unavailable on this box"
    gen = torch.Generator().manual_seed(100 + rank)
    for it in range(300):
        v = torch.randn(1 + it % 7, generator=gen, dtype=torch.float64).to(dev)
        a, b = v.clone(), v.clone()
        px.all_reduce_(a)
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
        assert torch.allclose(a, b, rtol=1e-14, atol=1e-14), (it, a, b)
    torch.cuda.synchronize()

    # ---- MultiBoxLoss512 on the config 2 shape, batch of 8 sharded over the ranks ----
    pri = PR.ssd512_canonical_priors()
    n_total = 4 * world
    locs, scores, bx, lb = synth.make_train_batch(pri, n_total, 81, 100, 77)  # identical on every rank (seeded)
    lo, hi = shard_range(n_total, rank, world)
    results = {}
    for mode in ("peer", "nccl", "full"):
        os.environ["SBOD_PEER_EXCHANGE"] = "0" if mode == "nccl" else "1"
        crit = MultiBoxLoss512(pri.to(dev), cfg(81))
        a, b = (0, n_total) if mode == "full" else (lo, hi)
        crit.process_group = None if mode == "full" else group
        l_d = locs[a:b].to(dev).requires_grad_(True)
        s_d = scores[a:b].to(dev).requires_grad_(True)
        loss = crit(l_d, s_d, [x.to(dev) for x in bx[a:b]], [x.to(dev) for x in lb[a:b]])
        loss.backward()
        st = crit.last["state"]
        assert (st.comm is not None) == (mode == "peer")
        results[mode] = (loss.item(), l_d.grad.clone(), s_d.grad.clone(), st.sums.clone())
    assert results["peer"][0] == results["nccl"][0], (results["peer"][0], results["nccl"][0])
    assert torch.equal(results["peer"][3], results["nccl"][3])
    assert torch.equal(results["peer"][1], results["nccl"][1]) and torch.equal(results["peer"][2], results["nccl"][2])
    full = results["full"]
    assert abs(results["peer"][0] - full[0]) <= 1e-6 * abs(full[0]), (results["peer"][0], full[0])
    assert torch.allclose(results["peer"][1], full[1][lo:hi], rtol=1e-5, atol=1e-9)
    assert torch.allclose(results["peer"][2], full[2][lo:hi], rtol=1e-5, atol=1e-9)
    # every rank holds the same scalar
    t = torch.tensor([results["peer"][0]], dtype=torch.float64, device=dev)
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    assert float(tmax) == float(tmin)

    # ---- the sharded step from a CUDA graph (the epoch lives in device memory) ----
    os.environ["SBOD_PEER_EXCHANGE"] = "1"
    import shape_based_object_detection_b200 as S
    crit = MultiBoxLoss512(pri.to(dev), cfg(81))
    crit.process_group = group
    l_d = locs[lo:hi].to(dev).requires_grad_(True)
    s_d = scores[lo:hi].to(dev).requires_grad_(True)
    gt = S.pack_ground_truth([x.to(dev) for x in bx[lo:hi]], [x.to(dev) for x in lb[lo:hi]], dev)
    out = torch.zeros((), device=dev)

    def step():
        l_d.grad = None
        s_d.grad = None
        loss = crit.forward_packed(l_d, s_d, gt)
        loss.backward()
        out.copy_(loss.detach())

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    assert out.item() == results["peer"][0], (out.item(), results["peer"][0])
    del g

    # ---- two criteria in flight on two streams: each on its own exchange lane. The ranks enqueue them in
    # OPPOSITE order, so a shared mailbox would pair the sums of different batches; with lanes both match. ----
    crit_a, crit_b = MultiBoxLoss512(pri.to(dev), cfg(81)), MultiBoxLoss512(pri.to(dev), cfg(81))
    crit_a.process_group = crit_b.process_group = group
    crit_b.exchange_lane = 1
    locs2, scores2, bx2, lb2 = synth.make_train_batch(pri, n_total, 81, 100, 78)
    batches = {"a": (locs, scores, bx, lb), "b": (locs2, scores2, bx2, lb2)}
    want = {}
    for name, (lo_, sc_, bx_, lb_) in batches.items():  # whole batch, one process
        ref = MultiBoxLoss512(pri.to(dev), cfg(81))
        want[name] = ref(lo_.to(dev), sc_.to(dev), [x.to(dev) for x in bx_], [x.to(dev) for x in lb_]).item()
    s_a, s_b = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    got = {}
    torch.cuda.synchronize()
    order = ("a", "b") if rank % 2 == 0 else ("b", "a")
    for rep in range(3):
        for name in order:
            lo_, sc_, bx_, lb_ = batches[name]
            crit_x, stream_x = (crit_a, s_a) if name == "a" else (crit_b, s_b)
            with torch.cuda.stream(stream_x):
                got[name] = crit_x(lo_[lo:hi].to(dev), sc_[lo:hi].to(dev), [x.to(dev) for x in bx_[lo:hi]],
                                   [x.to(dev) for x in lb_[lo:hi]])
        torch.cuda.synchronize()
        for name in ("a", "b"):
            assert abs(got[name].item() - want[name]) <= 1e-6 * abs(want[name]), (rep, name, got[name].item(), want[name])

    # ---- RefineDet (two criteria share the communicator) and FCOS (six sums) ----
    prr = PR.refinedet512_priors()
    gen = torch.Generator().manual_seed(5)
    bxr, lbr = synth.make_gt(n_total, 60, 4, gen, dense=True)
    P = prr.size(0)
    ts = [torch.randn((n_total, P, 4), generator=gen) * 0.1, torch.randn((n_total, P, 2), generator=gen),
          torch.randn((n_total, P, 4), generator=gen) * 0.1, torch.randn((n_total, P, 4), generator=gen)]
    vals = {}
    for mode in ("peer", "nccl", "full"):
        os.environ["SBOD_PEER_EXCHANGE"] = "0" if mode == "nccl" else "1"
        a, b = (0, n_total) if mode == "full" else (lo, hi)
        crit = RefineDetLoss(prr.to(dev), cfg(4))
        crit.process_group = None if mode == "full" else group
        loss = crit(*[t[a:b].to(dev) for t in ts], [x.to(dev) for x in bxr[a:b]], [x.to(dev) for x in lbr[a:b]])
        vals[mode] = loss.item()
    assert vals["peer"] == vals["nccl"], vals
    assert abs(vals["peer"] - vals["full"]) <= 1e-6 * abs(vals["full"]), vals

    locations = compute_location()
    Pf = sum(l.size(0) for l in locations)
    gen = torch.Generator().manual_seed(6)
    bxf, lbf = synth.make_gt(n_total, 12, 9, gen)
    tf = [torch.rand((n_total, Pf, 4), generator=gen) * 0.3 + 0.01, torch.randn((n_total, Pf, 9), generator=gen),
          torch.randn((n_total, Pf), generator=gen)]
    vals = {}
    for mode in ("peer", "nccl", "full"):
        os.environ["SBOD_PEER_EXCHANGE"] = "0" if mode == "nccl" else "1"
        a, b = (0, n_total) if mode == "full" else (lo, hi)
        crit = FCOSLoss([l.to(dev) for l in locations], cfg(9))
        crit.process_group = None if mode == "full" else group
        xs = [t[a:b].to(dev).requires_grad_(True) for t in tf]
        loss = crit(*xs, [x.to(dev) for x in bxf[a:b]], [x.to(dev) for x in lbf[a:b]])
        loss.backward()
        vals[mode] = (loss.item(), xs[1].grad.clone())
    assert vals["peer"][0] == vals["nccl"][0], (vals["peer"][0], vals["nccl"][0])
    assert abs(vals["peer"][0] - vals["full"][0]) <= 1e-6 * abs(vals["full"][0])
    assert torch.allclose(vals["peer"][1], vals["full"][1][lo:hi], rtol=1e-5, atol=1e-9)

    dist.barrier()
    if rank == 0:
        print("multi_gpu_check ok: world %d, loss %.6f (peer == nccl, full batch %.6f)" % (world, results["peer"][0], full[0]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
