"""Mirror of the VOC-style mAP of reference core/metrics/mAP.py (`get_map` :302-835, `voc_ap` :107-150), fed from
detection rows instead of text files.

The reference writes one txt file per image, re-reads and re-parses all of them once per class, and matches every
detection to the ground truth in a Python loop over JSON files on disk.  Here the rows the evaluators already hold
(CVPP_ROWS_VOC layout, image-sharded, compact) are matched on the device by `cvpp_voc_match` (one CTA per image,
the reference's double-precision "+1" overlap and `used` semantics), and the host is left with what is inherently
sequential and tiny: per class, a stable sort by the 6-character confidence the reference parses back from its own
files, two cumulative sums and the precision-envelope integral of `voc_ap`, in Python floats like the reference.

Scope: get_map's numbers (AP per class, mAP, precision / recall arrays, TP counts).  Plotting, the animation, the
log-average miss rate report and `get_coco_map` (pycocotools) are not part of the detection hot path.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from ... import ops


def voc_ap(rec: Sequence[float], prec: Sequence[float]):
    """Area under the monotone precision envelope at the recall change points (reference :107-150).
    Returns (ap, mrec, mpre) like the reference, without mutating the arguments."""
    mrec = [0.0] + [float(r) for r in rec] + [1.0]
    mpre = [0.0] + [float(p) for p in prec] + [0.0]
    for i in range(len(mpre) - 2, -1, -1):
        mpre[i] = max(mpre[i], mpre[i + 1])
    ap = 0.0
    for i in range(1, len(mrec)):
        if mrec[i] != mrec[i - 1]:
            ap += (mrec[i] - mrec[i - 1]) * mpre[i]
    return ap, mrec, mpre


def truncated_confidence(scores: np.ndarray) -> np.ndarray:
    """float(str(np.float32(s))[:6]): the confidence get_map actually sorts by - it reads back the 6 characters
    evaluate_on_voc wrote (yolo_v8.py:288-296, mAP.py:441)."""
    return np.array([float(str(np.float32(s))[:6]) for s in scores], dtype=np.float64)


def get_map_from_rows(MINOVERLAP: float, det_rows, det_counts: Sequence[int], gt_boxes, gt_classes, gt_difficult,
                      gt_counts: Sequence[int], class_names: Sequence[str], device=None) -> Dict:
    """VOC mAP of a whole data set.

    det_rows (N, 6) float32 VOC rows [cls, score, l, t, r, b] of all images back to back in FILE order (the sorted
    image-id order get_map globs, :336,413), det_counts rows per image, every (image, class) group in descending
    score order (what the evaluators emit); gt_boxes (G, 4) l,t,r,b, gt_classes (G,) class ids, gt_difficult (G,)
    0/1, gt_counts boxes per image.  numpy arrays or torch tensors (device tensors are used in place).

    Returns {"ap": {class_name: ap}, "map": mean over the classes that have non-difficult ground truth (:693),
             "rec" / "prec": {class_name: list}, "tp": {class_name: count}, "classes": sorted class names}.
    """
    if device is None:
        device = det_rows.device if isinstance(det_rows, torch.Tensor) and det_rows.is_cuda else torch.device("cuda")

    def dev_t(x, dtype):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(device=device, dtype=dtype).contiguous()

    rows_d = dev_t(det_rows, torch.float32).reshape(-1, 6)
    det_off = torch.tensor(np.concatenate([[0], np.cumsum(det_counts)]), dtype=torch.int32, device=device)
    gt_off = torch.tensor(np.concatenate([[0], np.cumsum(gt_counts)]), dtype=torch.int32, device=device)
    gb, gc, gd = dev_t(gt_boxes, torch.float32).reshape(-1, 4), dev_t(gt_classes, torch.int32), dev_t(gt_difficult, torch.int32)
    flag_d, _, _ = ops.voc_match(rows_d, det_off, gb, gc, gd, gt_off, float(MINOVERLAP))
    # one transfer back: the class / score columns and the verdicts
    flag = flag_d.cpu().numpy()
    rows_h = rows_d[:, :2].cpu().numpy()
    cls_h, score_h = rows_h[:, 0].astype(np.int64), rows_h[:, 1]
    gc_h, gd_h = gc.cpu().numpy(), gd.cpu().numpy()

    n_gt = np.bincount(gc_h[gd_h == 0], minlength=len(class_names))        # non-difficult boxes per class (:389-393)
    names = sorted(class_names[c] for c in range(len(class_names)) if n_gt[c] > 0)
    index_of = {n: i for i, n in enumerate(class_names)}
    out = {"ap": {}, "rec": {}, "prec": {}, "tp": {}, "classes": names}
    sum_ap = 0.0
    for name in names:
        c = index_of[name]
        sel = np.nonzero(cls_h == c)[0]                                    # file order, then line order
        conf = truncated_confidence(score_h[sel])
        order = np.argsort(-conf, kind="stable")                           # list.sort(reverse=True) is stable (:441)
        f = flag[sel][order]
        tp = np.cumsum(f == 1)
        fp = np.cumsum(f == 2)
        rec = [float(t) / max(int(n_gt[c]), 1) for t in tp]
        prec = [float(t) / max(int(p + t), 1) for t, p in zip(tp, fp)]
        ap, _, _ = voc_ap(rec, prec)
        sum_ap += ap
        out["ap"][name], out["rec"][name], out["prec"][name] = ap, rec, prec
        out["tp"][name] = int(tp[-1]) if len(tp) else 0
    out["map"] = sum_ap / len(names) if names else 0.0
    return out
