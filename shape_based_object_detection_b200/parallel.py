"""Multi-GPU rule of the path: images are independent, so the batch is sharded by
image, one process per GPU; the only data that crosses NVLink are the four batch sums
[sum loc, sum conf over positives, sum conf over mined negatives, n_pos] — one all-reduce(sum) per
step — after which every rank forms the same scalar and backward scales by the global 1/n_pos.
The CUDA path does this inside core.fused_loss (NCCL all-reduce of LossState.sums, then
sbod_loss_finalize); the functions here are the host-side statement of the same rule, used by the
world_size-2 gloo test."""
import torch


def shard_range(n_images, rank, world):
    """Contiguous, balanced [lo, hi) slice of the batch owned by `rank`."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_loss(sums, reg_kind, cls_kind, reg_weight, group=None):
    """sums: float64[4] partial sums of this rank. Returns float64[4]: total, conf, loc, n_pos — the
    arithmetic of finalize_loss in csrc/loss.cu."""
    if group is not None:
        import torch.distributed as dist
        sums = sums.clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    s_loc, s_pos, s_neg, n_pos = [float(v) for v in sums]
    loc = s_loc / (4.0 * n_pos) if reg_kind == 0 else s_loc / n_pos
    conf = (s_pos + s_neg) if cls_kind == 3 else (s_pos + s_neg) / n_pos
    return torch.tensor([conf + reg_weight * loc, conf, loc, n_pos], dtype=torch.float64)
