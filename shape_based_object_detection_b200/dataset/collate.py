"""Ground truth of a batch as ONE pinned buffer.

The reference collates a batch into python lists of per-image tensors (dataset/Datasets.py:58-86) and moves
them to the device one tensor at a time (train_anchor.py:266-268: 2N small H2D copies per step). The fused
loss wants CSR: boxes [T,4] f32, labels [T] i64, offsets [N+1] i32. `collate_fn` below is a drop-in for
Datasets.collate_fn (same five return values, same list semantics) whose `boxes` list also carries the batch
packed in one pinned host buffer: `boxes.packed.to(device)` is a single H2D copy, and every *Loss.forward of
this package accepts the result in place of the `boxes` list (`labels` is then ignored).
"""
import numpy as np
import torch


def _align(n, a=16):
    return (n + a - 1) // a * a


class PackedGT:
    """CSR ground truth of a batch inside one uint8 buffer: [offsets i32 | labels i64 | boxes f32], every
    section 16-byte aligned (the kernels read boxes as float4)."""

    def __init__(self, buf, n_images, total, gmax):
        self.buf, self.n_images, self.total, self.gmax = buf, int(n_images), int(total), int(gmax)
        t = max(self.total, 1)
        self._o_off = 0
        self._o_lab = _align(4 * (self.n_images + 1))
        self._o_box = self._o_lab + _align(8 * t)
        self.nbytes = self._o_box + 16 * t

    @staticmethod
    def layout_bytes(n_images, total):
        t = max(int(total), 1)
        return _align(4 * (n_images + 1)) + _align(8 * t) + 16 * t

    @classmethod
    def from_lists(cls, boxes, labels, pin=True):
        counts = [int(b.shape[0]) for b in boxes]
        n, total = len(counts), sum(counts)
        buf = torch.zeros(cls.layout_bytes(n, total), dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            buf = buf.pin_memory()
        gt = cls(buf, n, total, max(counts) if counts else 0)
        offs = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(counts, out=offs[1:])
        gt.offsets.copy_(torch.from_numpy(offs))
        if total:
            torch.cat([b.reshape(-1, 4).to(torch.float32) for b in boxes], 0, out=gt.boxes)
            torch.cat([l.reshape(-1).to(torch.int64) for l in labels], 0, out=gt.labels)
        return gt

    # ---- typed views of the buffer ----
    @property
    def offsets(self):
        return self.buf[self._o_off:self._o_off + 4 * (self.n_images + 1)].view(torch.int32)

    @property
    def labels(self):
        return self.buf[self._o_lab:self._o_lab + 8 * max(self.total, 1)].view(torch.int64)

    @property
    def boxes(self):
        return self.buf[self._o_box:self._o_box + 16 * max(self.total, 1)].view(torch.float32).view(-1, 4)

    @property
    def device(self):
        return self.buf.device

    def to(self, device, non_blocking=True):
        """ONE copy of the whole batch's ground truth."""
        return PackedGT(self.buf.to(device, non_blocking=non_blocking), self.n_images, self.total, self.gmax)

    def as_tuple(self):
        """(boxes, labels, offsets, gmax) as core.pack_ground_truth returns it."""
        return self.boxes, self.labels, self.offsets, self.gmax

    def lists(self):
        """The reference's two lists of per-image tensors (views, no copy)."""
        offs = self.offsets.tolist()
        return ([self.boxes[offs[i]:offs[i + 1]] for i in range(self.n_images)],
                [self.labels[offs[i]:offs[i + 1]] for i in range(self.n_images)])


class GTList(list):
    """The reference's list of per-image box tensors, plus `.packed` (PackedGT, pinned host memory)."""
    packed = None


def collate_fn(batch):
    """Drop-in for Datasets.collate_fn (dataset/Datasets.py:58-86): images stacked, boxes / labels / ids /
    difficulties as lists; `boxes.packed` additionally holds boxes + labels in one pinned CSR buffer."""
    images, boxes, labels, ids, difficulties = [], GTList(), [], [], []
    for b in batch:
        images.append(b[0])
        boxes.append(b[1])
        labels.append(b[2])
        ids.append(b[3])
        difficulties.append(b[4])
    images = torch.stack(images, dim=0)
    boxes.packed = PackedGT.from_lists(boxes, labels)
    return images, boxes, labels, ids, difficulties
