"""COCO-format result records from detect()'s per-image outputs — the loop at eval.py:185-213
(SURVEY §8f rank 4). Host-side formatting only: the xyxy -> xywh * image size conversion is done with
tensor ops on all detections at once (same fp32 operation order as the reference: w = x2 - x1, then
the four scalings) and the records are built from ONE device->host transfer instead of one
`float(tensor)` / `.tolist()` per box."""
import torch


def coco_format_results(det_boxes_batch, det_labels_batch, det_scores_batch, image_ids, image_sizes,
                        category_of_label):
    """det_*_batch: lists with one tensor per image, as returned by detect() ([n_i, 4] xyxy in [0, 1],
    [n_i] labels, [n_i] scores); image_ids: list of COCO image ids; image_sizes: list of (width, height);
    category_of_label: {label id -> COCO category id} (the reference looks it up per box through
    rev_coco_label_map and coco.getCatIds, eval.py:204-205). Returns the list of
    {'image_id', 'category_id', 'score', 'bbox': [x, y, w, h]} dicts in the reference's order."""
    assert len(det_boxes_batch) == len(det_labels_batch) == len(det_scores_batch) == len(image_ids) == len(image_sizes)
    counts = [int(b.size(0)) for b in det_boxes_batch]
    if sum(counts) == 0:
        return []
    boxes = torch.cat([b.reshape(-1, 4) for b in det_boxes_batch], 0).to(torch.float32)
    dev = boxes.device
    scale = torch.tensor([[float(w) * 1., float(h) * 1.] for (w, h) in image_sizes], dtype=torch.float32, device=dev)
    scale = scale.repeat_interleave(torch.tensor(counts, device=dev), dim=0)          # [D, 2] = (width, height)
    wh = boxes[:, 2:] - boxes[:, :2]                                                   # eval.py:194-195
    out = torch.cat([boxes[:, :2] * scale, wh * scale], 1)                             # eval.py:196-199
    bbox = out.cpu().tolist()
    scores = torch.cat([s.reshape(-1) for s in det_scores_batch], 0).to(torch.float32).cpu().tolist()
    labels = torch.cat([l.reshape(-1) for l in det_labels_batch], 0).cpu().tolist()
    results, k = [], 0
    for image_id, n in zip(image_ids, counts):
        for _ in range(n):
            results.append({'image_id': image_id, 'category_id': category_of_label[int(labels[k])],
                            'score': scores[k], 'bbox': bbox[k]})
            k += 1
    return results
