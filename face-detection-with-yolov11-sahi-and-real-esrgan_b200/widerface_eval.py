"""(f1) WIDER-FACE official-protocol AP with the IoU matrix on the GPU.

Drop-in for the maths of eval/eval_official_widerface.py:282-453: `bbox_overlaps` (the reference's only native
dependency, a Cython module of the external WiderFace-Evaluation repo) is the `fsd_bbox_overlaps_p1` kernel; the
greedy per-image matching is order dependent and stays a host loop, the 1000-threshold PR accumulation is vectorised."""
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
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def image_eval(pred, gt, ignore, iou_thresh=0.5, device="cuda:0"):
    p, g = pred.copy(), gt.copy()
    p[:, 2] += p[:, 0]
    p[:, 3] += p[:, 1]
    g[:, 2] += g[:, 0]
    g[:, 3] += g[:, 1]
    overlaps = bbox_overlaps(p[:, :4], g, device)
    best = overlaps.argmax(1)
    best_v = overlaps[np.arange(len(p)), best]
    recall = np.zeros(g.shape[0])
    proposal = np.ones(p.shape[0])
    pred_recall = np.zeros(p.shape[0])
    matched = 0
    for h in range(p.shape[0]):
        if best_v[h] >= iou_thresh:
            j = best[h]
            if ignore[j] == 0:
                if recall[j] == 1:
                    matched -= 1
                recall[j] = -1
                proposal[h] = -1
            elif recall[j] == 0:
                recall[j] = 1
                matched += 1
        pred_recall[h] = matched
    return pred_recall, proposal


def img_pr_info(thresh_num, pred_info, proposal_list, pred_recall):
    thresh = 1 - (np.arange(thresh_num) + 1) / thresh_num
    scores = pred_info[:, 4]
    # last index with score >= thresh (predictions are not assumed sorted: same rule as np.where(...)[-1])
    ge = scores[None, :] >= thresh[:, None]
    any_ge = ge.any(1)
    last = ge.shape[1] - 1 - np.argmax(ge[:, ::-1], axis=1)
    cum_valid = np.cumsum(proposal_list == 1)
    out = np.zeros((thresh_num, 2))
    out[any_ge, 0] = cum_valid[last[any_ge]]
    out[any_ge, 1] = pred_recall[last[any_ge]]
    return out


def evaluate_setting(preds, gts, keep_indices, thresh_num=1000, iou_thresh=0.5, device="cuda:0"):
    count_face = 0
    pr_curve = np.zeros((thresh_num, 2))
    for pred_info, gt_boxes, keep_index in zip(preds, gts, keep_indices):
        count_face += len(keep_index)
        if len(gt_boxes) == 0 or len(pred_info) == 0:
            continue
        ignore = np.zeros(gt_boxes.shape[0])
        if len(keep_index) != 0:
            ignore[np.asarray(keep_index) - 1] = 1
        pred_recall, proposal = image_eval(pred_info.astype(float), gt_boxes.astype(float), ignore, iou_thresh, device)
        pr_curve += img_pr_info(thresh_num, pred_info, proposal, pred_recall)
    out = np.zeros_like(pr_curve)
    nz = pr_curve[:, 0] != 0
    out[nz, 0] = pr_curve[nz, 1] / pr_curve[nz, 0]
    out[:, 1] = pr_curve[:, 1] / count_face
    return voc_ap(out[:, 1], out[:, 0]), out
