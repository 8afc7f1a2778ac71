"""Sub-category / difficulty evaluator of the reference's `eval/eval_dual.py` as array code.

Same numbers as `DualWiderFaceEvaluator.evaluate_single_set / calculate_average_precision / calculate_summary_metrics`
(eval/eval_dual.py:292-315, 334-433, 496-515): per image one IoU matrix (predictions x faces, plain xywh IoU of :272-290), the
greedy first-best matching in prediction order, "ignored" faces of other sub-categories, 11-point interpolated AP.
This is host-side metric bookkeeping (a few hundred boxes per image); the detections it consumes come from the device path.
"""
from __future__ import annotations

import numpy as np

SUBCATEGORIES = ["large_clear", "large_degraded", "medium_clear", "medium_degraded", "small_clear", "small_degraded"]
DIFFICULTY_MAPPING = {"easy": ["large_clear"], "medium": ["large_clear", "large_degraded", "medium_clear"], "hard": SUBCATEGORIES}


def iou_xywh_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """[n,4] x [m,4] xywh -> [n,m]; 0 where the boxes do not overlap or the union is not positive (eval_dual.py:272-290)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 4)
    ax2, ay2 = a[:, 0] + a[:, 2], a[:, 1] + a[:, 3]
    bx2, by2 = b[:, 0] + b[:, 2], b[:, 1] + b[:, 3]
    ix1 = np.maximum(a[:, None, 0], b[None, :, 0])
    iy1 = np.maximum(a[:, None, 1], b[None, :, 1])
    ix2 = np.minimum(ax2[:, None], bx2[None, :])
    iy2 = np.minimum(ay2[:, None], by2[None, :])
    inter = (ix2 - ix1) * (iy2 - iy1)
    union = (a[:, 2] * a[:, 3])[:, None] + (b[:, 2] * b[:, 3])[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.where(union > 0, inter / union, 0.0)
    iou[(ix2 < ix1) | (iy2 < iy1)] = 0.0
    return iou


def average_precision_11pt(confidence, is_tp, total_gt: int) -> float:
    confidence = np.asarray(confidence, dtype=np.float64)
    is_tp = np.asarray(is_tp, dtype=bool)
    if total_gt == 0 or confidence.size == 0:
        return 0.0
    order = np.argsort(-confidence, kind="stable")  # stable descending == list.sort(reverse=True) on equal keys
    tp = np.cumsum(is_tp[order])
    fp = np.cumsum(~is_tp[order])
    recalls = tp / total_gt
    precisions = tp / (tp + fp)
    ap = 0.0
    for t in np.arange(0.0, 1.1, 0.1):
        sel = recalls >= t
        ap += (np.max(precisions[sel]) if sel.any() else 0) / 11.0
    return float(ap)


def evaluate_single_set(subcategory_gt: dict, predictions: dict, category_name: str, valid_categories,
                        iou_threshold: float = 0.5, global_confidence: float = 0.25) -> dict:
    total_gt = false_negatives = 0
    confs, tps = [], []
    for img_path, gt in subcategory_gt.items():
        valid = list(set(i for cat in valid_categories for i in gt[cat]))
        if not valid:
            continue
        faces = np.asarray([f["bbox"] for f in gt["all_faces"]], dtype=np.float64).reshape(-1, 4)
        gt_boxes = faces[valid]
        total_gt += len(valid)
        vset = set(valid)
        ign_boxes = faces[[i for i in range(len(faces)) if i not in vset]]
        preds = predictions.get(img_path, [])
        matched = np.zeros(len(valid), dtype=bool)
        if preds:
            pb = np.asarray([p["bbox"] for p in preds], dtype=np.float64).reshape(-1, 4)
            iou = iou_xywh_matrix(pb, gt_boxes)
            best = iou.argmax(1)                      # first maximum == the reference's strict '>' scan
            best_iou = iou[np.arange(len(preds)), best]
            hits_ignored = (iou_xywh_matrix(pb, ign_boxes) >= iou_threshold).any(1) if len(ign_boxes) else np.zeros(len(preds), bool)
            for k, p in enumerate(preds):
                if best_iou[k] >= iou_threshold and best_iou[k] > 0 and not matched[best[k]]:
                    matched[best[k]] = True
                    confs.append(p["confidence"]); tps.append(True)
                elif not hits_ignored[k]:
                    confs.append(p["confidence"]); tps.append(False)
        false_negatives += int((~matched).sum())
    ap = average_precision_11pt(confs, tps, total_gt)
    confs_a, tps_a = np.asarray(confs, dtype=np.float64), np.asarray(tps, dtype=bool)
    keep = confs_a >= global_confidence
    n_keep, tp = int(keep.sum()), int((tps_a & keep).sum())
    precision = tp / n_keep if n_keep else 0
    recall = tp / total_gt if total_gt > 0 else 0
    f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) > 0 else 0
    return {"category": category_name, "total_gt": total_gt, "total_pred": n_keep, "true_positives": tp,
            "false_positives": n_keep - tp, "false_negatives": false_negatives, "precision": precision, "recall": recall,
            "f1_score": f1, "ap": ap}


def calculate_summary_metrics(sub, diff) -> dict:
    s = {"subcategory_map": float(np.mean([r["ap"] for r in sub]))}
    for key in ("large", "medium", "small", "clear", "degraded"):
        s[f"{key}_map"] = float(np.mean([r["ap"] for r in sub if key in r["category"]]))
    s["standard_map"] = float(np.mean([r["ap"] for r in diff]))
    for d in ("easy", "medium", "hard"):
        s[f"{d}_ap"] = next(r["ap"] for r in diff if r["category"] == d)
    return s


def evaluate_all(subcategory_gt: dict, predictions: dict, iou_threshold: float = 0.5, global_confidence: float = 0.25):
    """(sub-category results, easy/medium/hard results, summary) — DualWiderFaceEvaluator.run without printing / plotting."""
    sub = [evaluate_single_set(subcategory_gt, predictions, c, [c], iou_threshold, global_confidence) for c in SUBCATEGORIES]
    diff = [evaluate_single_set(subcategory_gt, predictions, d, DIFFICULTY_MAPPING[d], iou_threshold, global_confidence)
            for d in ("easy", "medium", "hard")]
    return sub, diff, calculate_summary_metrics(sub, diff)


def predictions_from_results(results, image_keys, scale: float = 1.0) -> dict:
    """{image: [{"bbox": xywh, "confidence": c}]} from PredictionResult objects (run_inference's conversion, eval_dual.py:247-266:
    xyxy -> xywh, coordinates divided by the enhancement scale)."""
    out = {}
    for key, res in zip(image_keys, results):
        rows = []
        for p in res.object_prediction_list:
            x1, y1, x2, y2 = p.bbox.to_xyxy()
            rows.append({"bbox": [c / scale for c in (x1, y1, x2 - x1, y2 - y1)] if scale != 1.0 else [x1, y1, x2 - x1, y2 - y1],
                         "confidence": float(p.score.value)})
        out[key] = rows
    return out
