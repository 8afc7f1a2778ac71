#!/usr/bin/env python
"""Golden vectors for the secondary evaluator, produced by the reference's OWN methods.

Imports `eval/eval_dual.py` unmodified from /root/reference (its un-vendored imports — matplotlib, seaborn, sahi,
utils.yolo_wrapper, utils.enhancer — are satisfied by empty in-memory shims: none of them is touched by the arithmetic),
builds a `DualWiderFaceEvaluator` WITHOUT running its constructor (which loads model files), feeds it seeded synthetic
ground truth and predictions through its prediction cache, and records what `calculate_iou`,
`calculate_average_precision`, `evaluate_single_set` and `calculate_summary_metrics` return.

    python tests/golden/make_golden_eval_dual.py      # writes tests/golden/eval_dual_outputs.json
"""
import importlib.util
import json
import os
import sys
import types
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SUBCATS = ["large_clear", "large_degraded", "medium_clear", "medium_degraded", "small_clear", "small_degraded"]


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_evaluator():
    _mod("matplotlib").__path__ = []
    _mod("matplotlib.pyplot")
    _mod("seaborn")
    _mod("sahi").__path__ = []
    _mod("sahi.predict", get_sliced_prediction=None)
    _mod("utils").__path__ = []
    _mod("utils.yolo_wrapper", YOLOv11PoseDetectionModel=None)
    _mod("utils.enhancer", FaceEnhancer=None)
    spec = importlib.util.spec_from_file_location("ref_eval_dual", os.path.join(REF, "eval", "eval_dual.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DualWiderFaceEvaluator


def synthetic_case(seed, n_images=6, faces_per_image=9):
    """Seeded GT (faces tagged with one sub-category each) and predictions (jittered faces, duplicates, false alarms)."""
    rng = np.random.default_rng(seed)
    gt, preds = {}, {}
    for i in range(n_images):
        key = f"{i}--Scene/img_{seed}_{i}.jpg"
        faces, cats = [], {c: [] for c in SUBCATS}
        for j in range(int(rng.integers(1, faces_per_image + 1))):
            w, h = float(rng.integers(8, 160)), float(rng.integers(8, 160))
            x, y = float(rng.integers(0, 900)), float(rng.integers(0, 600))
            faces.append({"bbox": [x, y, w, h]})
            cats[SUBCATS[int(rng.integers(0, 6))]].append(j)
        entry = {"all_faces": faces}
        entry.update(cats)
        gt[key] = entry
        rows = []
        for f in faces:
            if rng.random() < 0.8:
                jit = rng.normal(0, 0.12, 4) * np.array([f["bbox"][2], f["bbox"][3], f["bbox"][2], f["bbox"][3]])
                b = (np.array(f["bbox"]) + jit).round(2)
                rows.append({"bbox": [float(v) for v in b], "confidence": float(np.round(rng.uniform(0.02, 1.0), 3))})
                if rng.random() < 0.3:  # duplicate detection of the same face
                    rows.append({"bbox": [float(v) for v in (b + 1.5)], "confidence": float(np.round(rng.uniform(0.02, 1.0), 3))})
        for _ in range(int(rng.integers(0, 4))):  # false alarms; some share a confidence value (tie order matters)
            rows.append({"bbox": [float(rng.integers(0, 900)), float(rng.integers(0, 600)), 30.0, 40.0], "confidence": 0.5})
        order = rng.permutation(len(rows))
        preds[key] = [rows[k] for k in order]
    # one image without any face of some categories / without predictions
    lonely = {"all_faces": [{"bbox": [10.0, 10.0, 50.0, 60.0]}]}
    lonely.update({c: [] for c in SUBCATS})
    lonely["large_clear"] = [0]
    gt[f"empty--Scene/img_{seed}.jpg"] = lonely
    preds[f"empty--Scene/img_{seed}.jpg"] = []
    return gt, preds


def main():
    Evaluator = load_reference_evaluator()
    golden = {"iou_pairs": [], "ap_cases": [], "cases": []}
    ev = object.__new__(Evaluator)
    rng = np.random.default_rng(0)
    for _ in range(40):
        a = [float(v) for v in rng.integers(0, 60, 4)]
        b = [float(v) for v in rng.integers(0, 60, 4)]
        golden["iou_pairs"].append({"a": a, "b": b, "iou": float(ev.calculate_iou(a, b))})
    for n, total in ((0, 5), (7, 0), (1, 1), (12, 9), (40, 25), (40, 90)):
        dets = [{"confidence": float(np.round(rng.uniform(0, 1), 2)), "is_tp": bool(rng.random() < 0.6)} for _ in range(n)]
        golden["ap_cases"].append({"detections": dets, "total_gt": total,
                                   "ap": float(ev.calculate_average_precision([dict(d) for d in dets], total))})
    for seed in (1, 2, 3):
        gt, preds = synthetic_case(seed)
        ev = object.__new__(Evaluator)
        ev.subcategory_gt, ev.images_path = gt, Path("")
        ev.prediction_cache = {str(Path("") / k): v for k, v in preds.items()}
        ev.enhancement_stats = {"total_images": 0, "enhanced_images": 0, "skipped_images": 0}
        ev.use_enhancer, ev.iou_threshold, ev.global_confidence = False, 0.5, 0.25
        ev.temp_enh_dir = Path("/nonexistent")
        ev.subcategories, ev.difficulties = list(SUBCATS), ["easy", "medium", "hard"]
        mapping = {"easy": ["large_clear"], "medium": ["large_clear", "large_degraded", "medium_clear"], "hard": list(SUBCATS)}
        sub = [ev.evaluate_single_set("subcategory", c, [c]) for c in SUBCATS]
        diff = [ev.evaluate_single_set("difficulty", d, mapping[d]) for d in ev.difficulties]
        summary = ev.calculate_summary_metrics(sub, diff)
        tofloat = lambda r: {k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in r.items()}  # noqa: E731
        golden["cases"].append({"seed": seed, "gt": gt, "predictions": preds, "subcategory": [tofloat(r) for r in sub],
                                "difficulty": [tofloat(r) for r in diff], "summary": {k: float(v) for k, v in summary.items()},
                                "difficulty_of": {c: ev.map_subcategory_to_difficulty(c) for c in SUBCATS}})
    out = os.path.join(HERE, "eval_dual_outputs.json")
    with open(out, "w") as f:
        json.dump(golden, f, indent=0)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
