#!/usr/bin/env python
"""Golden vectors for the WIDER-FACE official-protocol evaluator, produced by the reference's OWN methods.

Imports `eval/eval_official_widerface.py` unmodified from /root/reference, builds an `OfficialWiderFaceEvaluator` WITHOUT
running its constructor (which loads model files), hands it seeded synthetic ground truth in the exact nested layout
`scipy.io.loadmat` gives for the official .mat files (the generator writes real .mat files with `savemat` and loads them
back through the reference's own `_load_official_ground_truth`), and records what `_voc_ap`, `_image_eval`, `_img_pr_info`,
`_dataset_pr_info` and `_evaluate_setting` (eval/eval_official_widerface.py:282-453) return.

The reference's un-vendored imports are satisfied by in-memory shims that the arithmetic never touches (matplotlib,
seaborn, sahi, utils.*), with ONE exception: `bbox.bbox_overlaps` is a Cython module of the external WiderFace-Evaluation
repository (not under /root/reference); it is bound to oracle.widerface_eval.bbox_overlaps, the restatement of its published
algorithm — that one function stays "parity unpinned".

    python tests/golden/make_golden_widerface_eval.py      # writes tests/golden/widerface_eval_outputs.json
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
from scipy.io import savemat

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_evaluator():
    from oracle import widerface_eval as oe

    _mod("matplotlib").__path__ = []
    _mod("matplotlib.pyplot")
    _mod("seaborn")
    _mod("bbox", bbox_overlaps=oe.bbox_overlaps)
    _mod("sahi").__path__ = []
    _mod("sahi.predict", get_sliced_prediction=None)
    _mod("utils").__path__ = []
    _mod("utils.yolo_wrapper", YOLOv11PoseDetectionModel=None)
    _mod("utils.enhancer", FaceEnhancer=None)
    spec = importlib.util.spec_from_file_location("ref_eval_official", os.path.join(REF, "eval", "eval_official_widerface.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.OfficialWiderFaceEvaluator


def synthetic_dataset(seed, n_events=3, images_per_event=5):
    """Events -> images -> GT boxes (xywh float) + per-setting keep lists (1-based), and predictions [n,5] xywh+score in
    the order the evaluator stores them (score-descending, as `_run_single_inference` sorts)."""
    rng = np.random.default_rng(seed)
    events = []
    for e in range(n_events):
        imgs = []
        for j in range(images_per_event):
            k = int(rng.integers(0, 10))
            gt = np.stack([rng.integers(0, 900, k), rng.integers(0, 600, k), rng.integers(6, 160, k), rng.integers(6, 200, k)], 1).astype(float) \
                if k else np.zeros((0, 4))
            keep = {"easy": [i + 1 for i in range(k) if gt[i, 3] > 50], "medium": [i + 1 for i in range(k) if gt[i, 3] > 20],
                    "hard": [i + 1 for i in range(k) if gt[i, 3] > 8]}
            det = []
            for g in gt:
                if rng.random() < 0.8:
                    b = g + rng.normal(0, 0.1, 4) * np.array([g[2], g[3], g[2], g[3]])
                    det.append([*np.round(b, 2), float(np.round(rng.uniform(0.02, 1.0), 3))])
                    if rng.random() < 0.3:  # duplicate detection of one face: only the first one counts
                        det.append([*np.round(b + 1.5, 2), float(np.round(rng.uniform(0.02, 1.0), 3))])
            for _ in range(int(rng.integers(0, 5))):  # false alarms, some with tied scores
                det.append([float(rng.integers(0, 900)), float(rng.integers(0, 600)), 30.0, 40.0, 0.5])
            det = np.array(sorted(det, key=lambda r: -r[4]), dtype=float).reshape(-1, 5)
            imgs.append(dict(name=f"{e}_Scene_{seed}_{j}", gt=gt, keep=keep, pred=det))
        events.append(dict(name=f"{e}--Scene{seed}", images=imgs))
    # an image with ground truth but no prediction, and one with predictions but no ground truth
    events[0]["images"][0]["pred"] = np.zeros((0, 5))
    events[1]["images"][1]["gt"] = np.zeros((0, 4))
    events[1]["images"][1]["keep"] = {"easy": [], "medium": [], "hard": []}
    return events


def write_mats(events, folder):
    def cell(items):
        c = np.empty((len(items), 1), dtype=object)
        for i, it in enumerate(items):
            c[i, 0] = it
        return c

    savemat(os.path.join(folder, "wider_face_val.mat"), {
        "face_bbx_list": cell([cell([im["gt"] for im in ev["images"]]) for ev in events]),
        "event_list": cell([np.array([ev["name"]]) for ev in events]),
        "file_list": cell([cell([np.array([im["name"]]) for im in ev["images"]]) for ev in events])})
    for setting in ("easy", "medium", "hard"):
        savemat(os.path.join(folder, f"wider_{setting}_val.mat"), {
            "gt_list": cell([cell([np.array(im["keep"][setting], dtype=np.int32).reshape(-1, 1) for im in ev["images"]]) for ev in events])})


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def main():
    Evaluator = load_reference_evaluator()
    golden = {"voc_ap": [], "image_eval": [], "cases": []}
    ev = object.__new__(Evaluator)
    ev.iou_threshold, ev.thresh_num = 0.5, 1000
    rng = np.random.default_rng(0)
    for n in (1, 5, 50, 1000):
        rec = np.sort(np.round(rng.uniform(0, 1, n), 3))
        prec = np.round(rng.uniform(0, 1, n), 3)
        golden["voc_ap"].append({"rec": rec.tolist(), "prec": prec.tolist(), "ap": float(ev._voc_ap(rec, prec))})
    for seed in (11, 12):
        events = synthetic_dataset(seed)
        with tempfile.TemporaryDirectory() as tmp:
            write_mats(events, tmp)
            ev = object.__new__(Evaluator)
            ev.gt_path, ev.settings, ev.iou_threshold, ev.thresh_num = Path(tmp), ["easy", "medium", "hard"], 0.5, 1000
            ev._load_official_ground_truth()  # the reference's own loadmat code path
        preds = {e["name"]: {im["name"]: im["pred"] for im in e["images"]} for e in events}
        case = {"seed": seed, "events": [{"name": e["name"], "images": [
            {"name": im["name"], "gt": im["gt"].tolist(), "keep": im["keep"], "pred": im["pred"].tolist()} for im in e["images"]]} for e in events],
            "settings": {}}
        for setting in ev.settings:
            ap, recall, propose = ev._evaluate_setting(setting, preds)
            entry = {"ap": float(ap), "recall_sha256": digest(recall), "propose_sha256": digest(propose)}
            if seed == 11 and setting == "hard":
                entry["recall"], entry["propose"] = [float(v) for v in recall], [float(v) for v in propose]
            case["settings"][setting] = entry
        golden["cases"].append(case)
        if seed == 11:  # the per-image building blocks on every image that reaches them
            for e in events:
                for im in e["images"]:
                    if len(im["gt"]) == 0 or len(im["pred"]) == 0:
                        continue
                    ignore = np.zeros(len(im["gt"]))
                    if im["keep"]["medium"]:
                        ignore[np.array(im["keep"]["medium"]) - 1] = 1
                    pr, pl = ev._image_eval(im["pred"].copy(), im["gt"].copy(), ignore)
                    info = ev._img_pr_info(im["pred"], pl, pr)
                    golden["image_eval"].append({"image": im["name"], "ignore": ignore.tolist(), "pred_recall": pr.tolist(),
                                                 "proposal_list": pl.tolist(), "pr_info_sha256": digest(info),
                                                 "pr_info_sum": [float(info[:, 0].sum()), float(info[:, 1].sum())]})
    curve = np.stack([np.arange(1000.0) % 7, np.arange(1000.0) % 5], 1)
    golden["dataset_pr_info"] = {"count_face": 37, "sha256": digest(ev._dataset_pr_info(curve, 37))}
    out = os.path.join(HERE, "widerface_eval_outputs.json")
    with open(out, "w") as f:
        json.dump(golden, f, indent=0)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
