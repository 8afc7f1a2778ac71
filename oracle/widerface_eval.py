"""Oracle for the WIDER-FACE official-protocol AP (test infrastructure).

Restates the reference's evaluator maths, eval/eval_official_widerface.py:282-453 (`_voc_ap`, `_image_eval`,
`_img_pr_info`, `_dataset_pr_info`, `_evaluate_setting`), and the Cython `bbox_overlaps` of the external
WiderFace-Evaluation repo it imports at :20-33 ("+1" pixel convention).  The reference omits the official score
normalisation step; so does this."""
from __future__ import annotations

import numpy as np


def bbox_overlaps(boxes: np.ndarray, query: np.ndarray) -> np.ndarray:
    n, k = boxes.shape[0], query.shape[0]
    out = np.zeros((n, k), dtype=np.float64)
    for j in range(k):
        qa = (query[j, 2] - query[j, 0] + 1) * (query[j, 3] - query[j, 1] + 1)
        for i in range(n):
            iw = min(boxes[i, 2], query[j, 2]) - max(boxes[i, 0], query[j, 0]) + 1
            if iw > 0:
                ih = min(boxes[i, 3], query[j, 3]) - max(boxes[i, 1], query[j, 1]) + 1
                if ih > 0:
                    ua = (boxes[i, 2] - boxes[i, 0] + 1) * (boxes[i, 3] - boxes[i, 1] + 1) + qa - iw * ih
                    out[i, j] = iw * ih / ua
    return out


def voc_ap(rec, prec):  # :282-300
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def image_eval(pred, gt, ignore, iou_thresh=0.5, overlaps_fn=bbox_overlaps):  # :302-349
    _pred, _gt = pred.copy(), gt.copy()
    pred_recall = np.zeros(_pred.shape[0])
    recall_list = np.zeros(_gt.shape[0])
    proposal_list = np.ones(_pred.shape[0])
    _pred[:, 2] += _pred[:, 0]
    _pred[:, 3] += _pred[:, 1]
    _gt[:, 2] += _gt[:, 0]
    _gt[:, 3] += _gt[:, 1]
    overlaps = overlaps_fn(_pred[:, :4], _gt)
    for h in range(_pred.shape[0]):
        gt_overlap = overlaps[h]
        max_overlap, max_idx = gt_overlap.max(), gt_overlap.argmax()
        if max_overlap >= iou_thresh:
            if ignore[max_idx] == 0:
                recall_list[max_idx] = -1
                proposal_list[h] = -1
            elif recall_list[max_idx] == 0:
                recall_list[max_idx] = 1
        pred_recall[h] = len(np.where(recall_list == 1)[0])
    return pred_recall, proposal_list


def img_pr_info(thresh_num, pred_info, proposal_list, pred_recall):  # :351-377
    pr_info = np.zeros((thresh_num, 2)).astype("float")
    for t in range(thresh_num):
        thresh = 1 - (t + 1) / thresh_num
        r_index = np.where(pred_info[:, 4] >= thresh)[0]
        if len(r_index) == 0:
            continue
        r_index = r_index[-1]
        p_index = np.where(proposal_list[: r_index + 1] == 1)[0]
        pr_info[t, 0] = len(p_index)
        pr_info[t, 1] = pred_recall[r_index]
    return pr_info


def dataset_pr_info(thresh_num, pr_curve, count_face):  # :379-396
    out = np.zeros((thresh_num, 2))
    for i in range(thresh_num):
        out[i, 0] = pr_curve[i, 1] / pr_curve[i, 0] if pr_curve[i, 0] != 0 else 0
        out[i, 1] = pr_curve[i, 1] / count_face
    return out


def evaluate_setting(preds, gts, keep_indices, thresh_num=1000, iou_thresh=0.5, overlaps_fn=bbox_overlaps):
    """preds[i]: [n,5] xywh+score (as the evaluator stores them), gts[i]: [k,4] xywh float, keep_indices[i]: 1-based
    indices of the ground-truth boxes that count in this setting (:398-453)."""
    count_face = 0
    pr_curve = np.zeros((thresh_num, 2), dtype=float)
    for pred_info, gt_boxes, keep_index in zip(preds, gts, keep_indices):
        count_face += len(keep_index)
        if len(gt_boxes) == 0 or len(pred_info) == 0:
            continue
        ignore = np.zeros(gt_boxes.shape[0])
        if len(keep_index) != 0:
            ignore[np.asarray(keep_index) - 1] = 1
        pred_recall, proposal_list = image_eval(pred_info.copy(), gt_boxes.copy(), ignore, iou_thresh, overlaps_fn)
        pr_curve += img_pr_info(thresh_num, pred_info, proposal_list, pred_recall)
    pr_curve = dataset_pr_info(thresh_num, pr_curve, count_face)
    return voc_ap(pr_curve[:, 1], pr_curve[:, 0]), pr_curve


def difficulty_keep_lists(gt_boxes, easy_px=50, medium_px=20):
    """Synthetic keep lists: easy = faces taller than 50 px, medium > 20 px, hard = all (1-based indices)."""
    h = gt_boxes[:, 3] if len(gt_boxes) else np.zeros(0)
    idx = np.arange(1, len(gt_boxes) + 1)
    return {"easy": idx[h > easy_px], "medium": idx[h > medium_px], "hard": idx}
