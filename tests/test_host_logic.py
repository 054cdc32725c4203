"""CPU suite: host-side logic — GT packing, synthetic generators, and the multi-GPU combination rule
(per-rank partial sums -> one all-reduce -> identical scalar on every rank), world_size 2 over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import LOSS_CASES, case_priors
from oracle import box_pipeline as O
from shape_based_object_detection_b200 import core, synth
from shape_based_object_detection_b200.parallel import combine_loss, shard_range


def test_pack_ground_truth_csr():
    boxes = [torch.rand(3, 4), torch.rand(0, 4), torch.rand(5, 4)]
    labels = [torch.tensor([1, 2, 3]), torch.zeros(0, dtype=torch.long), torch.tensor([4, 4, 1, 2, 9])]
    b, l, o, gmax = core.pack_ground_truth(boxes, labels, torch.device("cpu"))
    assert o.tolist() == [0, 3, 3, 8] and gmax == 5 and o.dtype == torch.int32
    assert torch.equal(b[3:], boxes[2]) and torch.equal(l[:3], labels[0]) and l.dtype == torch.int64
    b, l, o, gmax = core.pack_ground_truth([torch.rand(0, 4)], [torch.zeros(0, dtype=torch.long)], torch.device("cpu"))
    assert o.tolist() == [0, 0] and gmax == 0 and b.shape == (1, 4)


def test_synthetic_batches_are_seeded_and_in_range():
    pri = case_priors(LOSS_CASES["s512_sl1_ce"])
    a = synth.make_train_batch(pri, 2, 6, 8, 5)
    b = synth.make_train_batch(pri, 2, 6, 8, 5)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2][1], b[2][1])
    for bx, lb in zip(a[2], a[3]):
        assert 1 <= bx.size(0) <= 8 and float(bx.min()) >= 0 and float(bx.max()) <= 1
        assert int(lb.min()) >= 1 and int(lb.max()) <= 5
        assert bool((bx[:, 2:] > bx[:, :2]).all())


def test_shard_range_covers_batch():
    for n in (1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _oracle_sums(case, lo, hi):
    """[sum loc, sum conf_pos, sum conf_hardneg, n_pos] of images lo..hi, from the oracle's parts."""
    pri = case_priors(case)
    locs, scores, bx, lb = synth.make_train_batch(pri, case["N"], case["C"], case["gmax"], case["seed"])
    total, parts = O.multibox_loss(case["variant"], pri, locs[lo:hi], scores[lo:hi], bx[lo:hi], lb[lo:hi],
                                   want_parts=True)
    n_pos = float(parts["n_pos"].sum())
    loc_sum = float(parts["loc"]) * n_pos
    conf_sum = float(parts["conf"]) * n_pos
    return torch.tensor([loc_sum, conf_sum, 0.0, n_pos], dtype=torch.float64), float(total)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = LOSS_CASES["s512_sl1_ce"]
    lo, hi = shard_range(case["N"], rank, world)
    sums, _ = _oracle_sums(case, lo, hi)
    total = combine_loss(sums, reg_kind=1, cls_kind=0, reg_weight=1.0, group=dist.group.WORLD)
    out[rank] = float(total[0])
    dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_full_batch_loss():
    case = LOSS_CASES["s512_sl1_ce"]
    _, want = _oracle_sums(case, 0, case["N"])
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert abs(out[0] - out[1]) == 0.0
    assert abs(out[0] - want) <= 1e-6 * abs(want)


def test_bench_reference_arm_json_contract():
    """bench.py --impl reference (the CPU arm the driver runs beside ours) prints one JSON line with the
    contract's keys; it needs no GPU."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=repo)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "images/sec for assign+loss+NMS"
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in line, key
    assert line["unit"] == "images/s" and line["higher_is_better"] is True and line["value"] > 0
    assert "workload" in line["config"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["unit"] == line["unit"]
    # "reference" when oracle/_ref holds the reference's own files (build container / shipped to the GPU box)
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    # both arms must describe the workload with the same `config` object
    sys.path.insert(0, repo)
    import bench
    assert line["config"] == bench.config_entry(2)
    for cid in bench.CONFIGS:
        assert set(bench.config_entry(cid)) == {"workload", "config_id", "images_per_gpu", "l2"}


def test_reference_runner_port_equals_reference_when_present():
    """oracle/ref_runner.py: the reference's own code (oracle/_ref, when the build placed it there) and the
    oracle port give the same loss on the same small batch, for every config the reference can run."""
    from oracle import ref_runner as RR
    from shape_based_object_detection_b200 import priors as PR, synth
    if not RR.reference_available():
        pytest.skip("oracle/_ref not present (the reference lives in the build container only)")
    for cid, table, Cn, G in ((1, "ssd300", 21, 20), (2, "ssd512", 21, 30), (3, "retinanet", 21, 30)):
        pri = PR.PRIOR_TABLES[table]()[::3].contiguous()
        locs, scores, bx, lb = synth.make_train_batch(pri, 2, Cn, G, 50 + cid)
        a = RR.Runner(cid, pri, Cn, use_reference=True).train((locs, scores), bx, lb)
        b = RR.Runner(cid, pri, Cn, use_reference=False).train((locs, scores), bx, lb)
        assert abs(a - b) <= 1e-6 * abs(a), (cid, a, b)


def test_coco_format_results_matches_the_reference_loop():
    """eval.py:185-213 restated literally (per image, per box) against the batched formatter."""
    from shape_based_object_detection_b200.eval_results import coco_format_results
    g = torch.Generator().manual_seed(21)
    counts = [5, 0, 17, 1]
    boxes = [torch.rand((n, 4), generator=g) for n in counts]
    labels = [torch.randint(1, 81, (n,), generator=g) for n in counts]
    scores = [torch.rand((n,), generator=g) for n in counts]
    ids = [139, 285, 632, 724]
    sizes = [(640, 426), (586, 640), (640, 483), (375, 500)]
    cat = {i: 1000 + i for i in range(1, 81)}
    want = []
    for j in range(len(ids)):
        width, height = sizes[j][0] * 1., sizes[j][1] * 1.
        bb = boxes[j].clone()
        bb[:, 2] -= bb[:, 0]
        bb[:, 3] -= bb[:, 1]
        bb[:, 0] *= width
        bb[:, 2] *= width
        bb[:, 1] *= height
        bb[:, 3] *= height
        for k in range(bb.size(0)):
            want.append({'image_id': ids[j], 'category_id': cat[int(labels[j][k])], 'score': float(scores[j][k]),
                         'bbox': bb[k, :].tolist()})
    got = coco_format_results(boxes, labels, scores, ids, sizes, cat)
    assert got == want
    assert coco_format_results([torch.zeros((0, 4))], [torch.zeros((0,), dtype=torch.long)], [torch.zeros((0,))],
                               [1], [(10, 10)], cat) == []


def test_collate_fn_packs_the_batch_into_one_csr_buffer():
    """dataset.collate.collate_fn == the reference's Datasets.collate_fn (Datasets.py:58-86)
    plus the whole batch's ground truth in one (pinned) CSR buffer: one H2D copy instead of 2N."""
    from shape_based_object_detection_b200.dataset.collate import PackedGT, collate_fn
    g = torch.Generator().manual_seed(4)
    counts = [3, 0, 7, 1]
    batch = [(torch.rand((3, 8, 8), generator=g), torch.rand((n, 4), generator=g),
              torch.randint(1, 21, (n,), generator=g), 100 + i, torch.zeros(n, dtype=torch.uint8))
             for i, n in enumerate(counts)]
    images, boxes, labels, ids, diffs = collate_fn(batch)
    assert images.shape == (4, 3, 8, 8) and ids == [100, 101, 102, 103] and len(diffs) == 4
    assert isinstance(boxes, list) and all(torch.equal(a, b[1]) for a, b in zip(boxes, batch))  # reference semantics
    p = boxes.packed
    assert p.offsets.tolist() == [0, 3, 3, 10, 11] and p.gmax == 7 and p.total == 11
    assert torch.equal(p.boxes, torch.cat([b[1] for b in batch])) and p.boxes.data_ptr() % 16 == 0
    assert torch.equal(p.labels, torch.cat([b[2] for b in batch]))
    bl, ll = p.lists()
    assert all(torch.equal(a, b[1]) for a, b in zip(bl, batch)) and all(torch.equal(a, b[2]) for a, b in zip(ll, batch))
    q = p.to("cpu")  # the single copy
    assert torch.equal(q.boxes, p.boxes) and q.as_tuple()[3] == 7
    # same CSR as the device-side packer of the fused loss
    b2, l2, o2, gmax = core.pack_ground_truth([b[1] for b in batch], [b[2] for b in batch], torch.device("cpu"))
    assert torch.equal(b2, p.boxes) and torch.equal(l2, p.labels) and torch.equal(o2, p.offsets) and gmax == p.gmax
    empty = PackedGT.from_lists([torch.zeros((0, 4))], [torch.zeros((0,), dtype=torch.long)], pin=False)
    assert empty.offsets.tolist() == [0, 0] and empty.boxes.shape == (1, 4)
