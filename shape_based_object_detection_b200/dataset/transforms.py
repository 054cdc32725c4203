"""Box format converters of dataset/transforms.py:26-83, as CUDA kernels (differentiable with respect to
the boxes, like the reference's torch expressions)."""
import random

import torch

from .. import _boxops as B
from .. import _lib as L


def xy_to_cxcy(xy):
    """(x_min, y_min, x_max, y_max) -> (c_x, c_y, w, h); transforms.py:26-34."""
    return B.convert(xy, L.BOX_XY_TO_CXCY)


def cxcy_to_xy(cxcy):
    """(c_x, c_y, w, h) -> (x_min, y_min, x_max, y_max); transforms.py:37-45."""
    return B.convert(cxcy, L.BOX_CXCY_TO_XY)


def cxcy_to_gcxgcy(cxcy, priors_cxcy):
    """Encode centre-size boxes w.r.t. priors (variances 10 and 5); transforms.py:48-66."""
    return B.encode(cxcy, priors_cxcy, L.CODEC_TRANSFORMS, 0.1, 0.2)


def gcxgcy_to_cxcy(gcxgcy, priors_cxcy):
    """Decode model offsets into centre-size boxes; transforms.py:69-83."""
    return B.decode(gcxgcy, priors_cxcy, L.CODEC_TRANSFORMS, 0.1, 0.2)


def random_crop(image, boxes, labels, max_trials=50):
    """dataset/transforms.py:124-205 for CUDA tensors (the CPU user of find_jaccard_overlap).
    Same sampling, same python `random` stream, same result: the reference tries crops one at a time and calls
    find_jaccard_overlap(crop, boxes) + `.item()` per trial; here the (up to 50) trial crops of a round are
    drawn first, their overlaps with the boxes come from ONE dense-IoU kernel launch and one device->host read,
    and the random generator is put back to where the reference would have left it (right after the accepted
    trial), so whatever draws next sees the same stream.
    image (3, H, W), boxes (n, 4) boundary coordinates, labels (n) -> (new_image, new_boxes, new_labels)."""
    from ..metrics import find_jaccard_overlap
    L.need_cuda(boxes)
    original_h, original_w = image.size(1), image.size(2)
    while True:
        min_overlap = random.choice([0., .1, .3, .5, .7, .9, None])  # 'None' refers to no cropping
        if min_overlap is None:
            return image, boxes, labels
        crops, states = [], []
        for _ in range(max_trials):
            min_scale = 0.3
            scale_h = random.uniform(min_scale, 1)
            scale_w = random.uniform(min_scale, 1)
            new_h, new_w = int(scale_h * original_h), int(scale_w * original_w)
            aspect_ratio = new_h / new_w
            if not 0.5 < aspect_ratio < 2:
                continue  # (the reference draws nothing more for this trial)
            left = random.randint(0, original_w - new_w)
            top = random.randint(0, original_h - new_h)
            crops.append([left, top, left + new_w, top + new_h])
            states.append(random.getstate())
        if not crops:
            continue
        crop_t = torch.tensor(crops, dtype=torch.float32, device=boxes.device)
        overlap = find_jaccard_overlap(crop_t, boxes)  # (trials, n_objects): one launch for the whole round
        centers = (boxes[:, :2] + boxes[:, 2:]) / 2.
        inside = ((centers[None, :, 0] > crop_t[:, None, 0]) & (centers[None, :, 0] < crop_t[:, None, 2]) &
                  (centers[None, :, 1] > crop_t[:, None, 1]) & (centers[None, :, 1] < crop_t[:, None, 3]))
        ok = (~(overlap.max(dim=1)[0] < min_overlap)) & inside.any(dim=1)
        good = ok.nonzero().flatten()
        if good.numel() == 0:
            continue  # all trials failed: the generator is where the reference's would be
        t = int(good[0])
        random.setstate(states[t])
        left, top, right, bottom = crops[t]
        new_image = image[:, top:bottom, left:right]
        keep = inside[t]
        new_boxes = boxes[keep, :].clone()
        new_labels = labels[keep]
        crop = crop_t[t]
        new_boxes[:, :2] = torch.max(new_boxes[:, :2], crop[:2])
        new_boxes[:, :2] -= crop[:2]
        new_boxes[:, 2:] = torch.min(new_boxes[:, 2:], crop[2:])
        new_boxes[:, 2:] -= crop[:2]
        return new_image, new_boxes, new_labels
