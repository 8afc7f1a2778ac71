"""TEST INFRASTRUCTURE — CPU restatement of the secondary evaluator's arithmetic.

Follows `eval/eval_dual.py` of the reference, loop for loop:
  calculate_iou                  eval/eval_dual.py:272-290   xywh boxes, plain (no "+1") IoU, 0 when the boxes do not overlap
  calculate_average_precision    eval/eval_dual.py:292-315   11-point interpolated AP over a confidence-sorted TP/FP list
  map_subcategory_to_difficulty  eval/eval_dual.py:317-332
  evaluate_single_set            eval/eval_dual.py:334-433   greedy first-best matching in PREDICTION order, "ignored" faces
  calculate_summary_metrics      eval/eval_dual.py:496-515
Pinned by tests/golden/eval_dual_outputs.json, produced by the reference's own methods (tests/golden/make_golden_eval_dual.py).
Only tests/ may import this module.
"""
import numpy as np

SUBCATEGORIES = ["large_clear", "large_degraded", "medium_clear", "medium_degraded", "small_clear", "small_degraded"]
DIFFICULTY_MAPPING = {"easy": ["large_clear"], "medium": ["large_clear", "large_degraded", "medium_clear"], "hard": SUBCATEGORIES}


def calculate_iou(box1, box2):
    x1, y1, w1, h1 = box1
    x2, y2, w2, h2 = box2
    ix1, iy1 = max(x1, x2), max(y1, y2)
    ix2, iy2 = min(x1 + w1, x2 + w2), min(y1 + h1, y2 + h2)
    if ix2 < ix1 or iy2 < iy1:
        return 0.0
    inter = (ix2 - ix1) * (iy2 - iy1)
    union = (w1 * h1) + (w2 * h2) - inter
    return inter / union if union > 0 else 0.0


def calculate_average_precision(all_detections, total_gt):
    if total_gt == 0 or not all_detections:
        return 0.0
    dets = sorted(all_detections, key=lambda d: d["confidence"], reverse=True)  # list.sort is stable: ties keep list order
    tp = np.cumsum([d["is_tp"] for d in dets])
    fp = np.cumsum([not d["is_tp"] for d in dets])
    recalls = tp / total_gt
    precisions = tp / (tp + fp)
    ap = 0.0
    for t in np.arange(0.0, 1.1, 0.1):
        p = 0 if np.sum(recalls >= t) == 0 else np.max(precisions[recalls >= t])
        ap += p / 11.0
    return ap


def map_subcategory_to_difficulty(category):
    out = []
    if category == "large_clear":
        out.append("easy")
    if category in ("large_clear", "large_degraded", "medium_clear"):
        out.append("medium")
    out.append("hard")
    return out


def evaluate_single_set(subcategory_gt, predictions, category_name, valid_categories, iou_threshold=0.5, global_confidence=0.25):
    """subcategory_gt: {image: {"all_faces": [{"bbox": xywh}, ...], <subcategory>: [face indices], ...}} in file order;
    predictions: {image: [{"bbox": xywh, "confidence": c}, ...]} in the detector's output order."""
    total_gt, all_detections, false_negatives = 0, [], 0
    for img_path, gt_data in subcategory_gt.items():
        valid = []
        for cat in valid_categories:
            valid.extend(gt_data[cat])
        valid = list(set(valid))
        if not valid:
            continue
        faces = gt_data["all_faces"]
        gt_faces = [faces[i] for i in valid]
        total_gt += len(gt_faces)
        ignored = [faces[i] for i in range(len(faces)) if i not in valid]
        matched = [False] * len(gt_faces)
        for pred in predictions.get(img_path, []):
            best_iou, best_idx, is_ignored = 0, -1, False
            for gi, g in enumerate(gt_faces):
                iou = calculate_iou(pred["bbox"], g["bbox"])
                if iou > best_iou:
                    best_iou, best_idx = iou, gi
            if best_iou >= iou_threshold and best_idx != -1 and not matched[best_idx]:
                matched[best_idx] = True
                is_tp = True
            else:
                for g in ignored:
                    if calculate_iou(pred["bbox"], g["bbox"]) >= iou_threshold:
                        is_ignored = True
                        break
                is_tp = False
            if not is_ignored:
                all_detections.append({"confidence": pred["confidence"], "is_tp": is_tp})
        false_negatives += sum(1 for m in matched if not m)
    ap = calculate_average_precision(all_detections, total_gt)
    kept = [d for d in all_detections if d["confidence"] >= global_confidence]
    tp = sum(1 for d in kept if d["is_tp"])
    fp = len(kept) - tp
    precision = tp / len(kept) if kept else 0
    recall = tp / total_gt if total_gt > 0 else 0
    f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) > 0 else 0
    return {"category": category_name, "total_gt": total_gt, "total_pred": len(kept), "true_positives": tp, "false_positives": fp,
            "false_negatives": false_negatives, "precision": precision, "recall": recall, "f1_score": f1, "ap": ap}


def calculate_summary_metrics(sub, diff):
    s = {"subcategory_map": np.mean([r["ap"] for r in sub])}
    for key in ("large", "medium", "small", "clear", "degraded"):
        s[f"{key}_map"] = np.mean([r["ap"] for r in sub if key in r["category"]])
    s["standard_map"] = np.mean([r["ap"] for r in diff])
    for d in ("easy", "medium", "hard"):
        s[f"{d}_ap"] = next(r["ap"] for r in diff if r["category"] == d)
    return s


def evaluate_all(subcategory_gt, predictions, iou_threshold=0.5, global_confidence=0.25):
    sub = [evaluate_single_set(subcategory_gt, predictions, c, [c], iou_threshold, global_confidence) for c in SUBCATEGORIES]
    diff = [evaluate_single_set(subcategory_gt, predictions, d, DIFFICULTY_MAPPING[d], iou_threshold, global_confidence)
            for d in ("easy", "medium", "hard")]
    return sub, diff, calculate_summary_metrics(sub, diff)
