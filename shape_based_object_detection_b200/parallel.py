"""Multi-GPU rule of the path: images are independent, so the batch is sharded by
image, one process per GPU; the only data that crosses NVLink are the four batch sums
[sum loc, sum conf over positives, sum conf over mined negatives, n_pos] — one all-reduce(sum) per
step — after which every rank forms the same scalar and backward scales by the global 1/n_pos.
The CUDA path does this inside core.fused_loss (NCCL all-reduce of LossState.sums, then
sbod_loss_finalize); the functions here are the host-side statement of the same rule, used by the
world_size-2 gloo test."""
import ctypes as C
import os
import socket
import warnings

import torch


# ---------------------------------------------------------------------------------------------
# In-kernel exchange of the loss sums (csrc/comm.cuh): mailboxes in every GPU's HBM, mapped into the
# processes of the node with CUDA IPC. One communicator per (process group, device, lane), created on first use.
# SBOD_PEER_EXCHANGE=0 keeps the NCCL all-reduce (the checked reference path of the exchange).
# ---------------------------------------------------------------------------------------------
_comms = {}


class PeerExchange:
    """Owns one sbod communicator. `ptr` goes into sbod_loss_desc.comm / sbod_fcos_desc.comm."""

    def __init__(self, group, device):
        import torch.distributed as dist
        from . import _lib as L
        lib = L.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = int(lib.sbod_comm_handle_bytes())
        handle = (C.c_ubyte * nbytes)()
        self._comm = C.c_void_p()
        ok = 1
        try:
            with torch.cuda.device(device):
                L.check(lib.sbod_comm_create(self.rank, self.world, C.byref(self._comm), handle))
        except L.SbodError:
            ok = 0
        # handles (and host names: CUDA IPC only reaches the GPUs of this node) of every rank
        mine = torch.tensor(list(bytes(handle)) + list(socket.gethostname().encode()[:64].ljust(64, b"\0")),
                            dtype=torch.uint8, device=device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        rows = torch.stack(gathered).cpu()
        same_host = bool((rows[:, nbytes:] == rows[0, nbytes:]).all())
        if ok and same_host:
            try:
                with torch.cuda.device(device):
                    L.check(lib.sbod_comm_connect(self._comm, rows[:, :nbytes].contiguous().numpy().tobytes()))
            except L.SbodError:
                ok = 0
        flag = torch.tensor([1 if (ok and same_host) else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)  # all ranks or none
        self.ok = bool(int(flag.item()))
        self.ptr = lib.sbod_comm_device_ptr(self._comm) if self.ok else None
        self._lib = lib

    def all_reduce_(self, sums):
        """Stand-alone in-place sum of a float64 tensor of <= 7 values through the mailboxes (tests)."""
        from . import _lib as L
        with torch.cuda.device(sums.device):
            L.check(self._lib.sbod_comm_allreduce(self._comm, L.ptr(sums), int(sums.numel()), L.stream_ptr()))
        return sums


def peer_exchange(group, device, lane=0):
    """The PeerExchange of (group, device, lane), or None when disabled / unavailable (then the caller all-reduces
    with torch.distributed, i.e. NCCL). Calls on one communicator must come in the same order on every rank: criteria
    that run concurrently on different streams (micro-batches in flight on two streams) take different lanes
    (`criterion.exchange_lane`), each with its own mailboxes and epoch counter. A communicator is created
    collectively on first use - not inside a CUDA graph capture."""
    if group is None or os.environ.get("SBOD_PEER_EXCHANGE", "1") == "0":
        return None
    key = (id(group), device.index, int(lane))
    if key not in _comms:
        px = PeerExchange(group, device)
        if not px.ok:
            warnings.warn("sbod: NVLink peer exchange unavailable (CUDA IPC between the ranks failed or the ranks "
                          "span several hosts): the loss sums are all-reduced with torch.distributed instead")
        _comms[key] = px
    px = _comms[key]
    return px if px.ok else None


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
