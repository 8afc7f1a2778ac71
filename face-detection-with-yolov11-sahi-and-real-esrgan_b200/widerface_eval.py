"""(f1) WIDER-FACE official-protocol AP on the GPU.

Drop-in for the maths of eval/eval_official_widerface.py:282-453.  `evaluate_setting` runs the whole data set of one
setting in ONE kernel launch (`fsd_widerface_pr_curve`): the "+1" IoU of the external Cython `bbox_overlaps`, the
order-dependent greedy matching of `_image_eval`, the 1000-threshold `_img_pr_info` and the accumulation over images all
happen on the device; only the [thresh_num, 2] curve comes back, and `_dataset_pr_info` / `_voc_ap` (2000 numbers) finish on
the host.  `image_eval` / `bbox_overlaps` expose the per-image pieces through the same kernels.  No CPU fallback: without
the CUDA library these raise."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def bbox_overlaps(boxes: np.ndarray, query: np.ndarray, device="cuda:0") -> np.ndarray:
    if boxes.shape[0] == 0 or query.shape[0] == 0:
        return np.zeros((boxes.shape[0], query.shape[0]), dtype=np.float64)
    b = torch.from_numpy(np.ascontiguousarray(boxes[:, :4], dtype=np.float64)).to(device)
    q = torch.from_numpy(np.ascontiguousarray(query[:, :4], dtype=np.float64)).to(device)
    return ops.bbox_overlaps_p1(b, q).cpu().numpy()


def voc_ap(rec, prec):
    """:282-300 (the precision envelope as a reversed running maximum)."""
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def _thresholds(thresh_num: int) -> np.ndarray:
    return np.array([1 - (t + 1) / thresh_num for t in range(thresh_num)], dtype=np.float64)  # python floats, as :364


def _launch(preds, gts, evaluates, thresh_num, iou_thresh, device):
    dev = torch.device(device)
    p_off = np.zeros(len(preds) + 1, dtype=np.int32)
    g_off = np.zeros(len(gts) + 1, dtype=np.int32)
    p_off[1:] = np.cumsum([len(p) for p in preds])
    g_off[1:] = np.cumsum([len(g) for g in gts])
    cat = lambda xs, w, dt: (np.concatenate([np.asarray(x, dtype=dt).reshape(-1, w) for x in xs]) if xs else np.zeros((0, w), dt))  # noqa: E731
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    pr, rec, prop = ops.widerface_pr_curve(up(cat(preds, 5, np.float64)), up(p_off), up(cat(gts, 4, np.float64)), up(g_off),
                                           up(cat(evaluates, 1, np.int32).reshape(-1)), up(_thresholds(thresh_num)), iou_thresh)
    return pr.cpu().numpy(), rec.cpu().numpy(), prop.cpu().numpy(), p_off


def image_eval(pred, gt, ignore, iou_thresh=0.5, device="cuda:0"):
    """:302-349 for one image: (pred_recall, proposal_list)."""
    _, rec, prop, _ = _launch([pred[:, :5]], [gt[:, :4]], [np.asarray(ignore) != 0], 1, iou_thresh, device)
    return rec, prop


def img_pr_info(thresh_num, pred_info, proposal_list, pred_recall):
    """:351-377 for given matching results (host helper kept for API parity; `evaluate_setting` does this on the device)."""
    thresh = _thresholds(thresh_num)
    ge = pred_info[:, 4][None, :] >= thresh[:, None]
    any_ge = ge.any(1)
    last = ge.shape[1] - 1 - np.argmax(ge[:, ::-1], axis=1)  # np.where(...)[0][-1]: predictions need not be sorted
    cum_valid = np.cumsum(proposal_list == 1)
    out = np.zeros((thresh_num, 2))
    out[any_ge, 0] = cum_valid[last[any_ge]]
    out[any_ge, 1] = pred_recall[last[any_ge]]
    return out


def evaluate_setting(preds, gts, keep_indices, thresh_num=1000, iou_thresh=0.5, device="cuda:0"):
    """:398-453.  preds[i] [n,5] xywh+score, gts[i] [k,4] xywh, keep_indices[i] 1-based indices of the ground-truth boxes
    that count in this setting.  Returns (AP, curve [thresh_num, 2] = precision, recall)."""
    count_face = 0
    P, Gt, Ev = [], [], []
    for pred_info, gt_boxes, keep_index in zip(preds, gts, keep_indices):
        count_face += len(keep_index)
        if len(gt_boxes) == 0 or len(pred_info) == 0:
            continue  # :428-429
        ev = np.zeros(len(gt_boxes), dtype=np.int32)
        if len(keep_index) != 0:
            ev[np.asarray(keep_index).reshape(-1) - 1] = 1
        P.append(np.asarray(pred_info, dtype=np.float64)[:, :5])
        Gt.append(np.asarray(gt_boxes, dtype=np.float64)[:, :4])
        Ev.append(ev)
    if P:
        pr_curve = _launch(P, Gt, Ev, thresh_num, iou_thresh, device)[0]
    else:
        pr_curve = np.zeros((thresh_num, 2))
    out = np.zeros_like(pr_curve)  # _dataset_pr_info :379-396
    nz = pr_curve[:, 0] != 0
    out[nz, 0] = pr_curve[nz, 1] / pr_curve[nz, 0]
    out[:, 1] = pr_curve[:, 1] / count_face
    return voc_ap(out[:, 1], out[:, 0]), out
