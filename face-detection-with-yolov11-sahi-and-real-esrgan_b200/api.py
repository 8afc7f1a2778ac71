"""Batch entry point of the fused path: many same-sized images in one call (what a serving loop or the evaluator's
dataset loop would use instead of calling get_sliced_prediction image by image)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import ops
from .sahi_api.prediction import ObjectPrediction, PredictionResult


def get_sliced_prediction_batch(images: Sequence, detection_model, slice_height: int, slice_width: int,
                                overlap_height_ratio: float = 0.2, overlap_width_ratio: float = 0.2,
                                perform_standard_pred: bool = True, postprocess_type: str = "GREEDYNMM",
                                postprocess_match_metric: str = "IOS", postprocess_match_threshold: float = 0.5,
                                postprocess_class_agnostic: bool = False, as_objects: bool = True, pool=None):
    """images: HWC uint8 arrays / CPU tensors (pinned for async H2D) of ONE common size.  Returns a list of
    PredictionResult (as_objects=True) or the raw engine.DetectionBatch."""
    eng = detection_model.engine()
    eng.truncate = True
    h, w = images[0].shape[:2]
    if pool is None or (pool.n, pool.h, pool.w) != (len(images), h, w):
        pool = ops.ImagePool(len(images), h, w, eng.device)
    for i, im in enumerate(images):
        pool.upload(i, im, non_blocking=True)
    batch = eng.detect(pool, slice_height, slice_width, overlap_height_ratio, overlap_width_ratio,
                       perform_standard_pred, postprocess_type, postprocess_match_metric,
                       postprocess_match_threshold, postprocess_class_agnostic)
    if not as_objects:
        return batch
    out: List[PredictionResult] = []
    for i in range(len(images)):
        boxes, scores, kpts, has_k = batch.image(i)
        preds = []
        for j in range(len(boxes)):
            op = ObjectPrediction(bbox=[int(v) for v in boxes[j]], category_id=0, category_name="face",
                                  score=float(scores[j]), shift_amount=[0, 0], full_shape=None)
            if has_k[j]:
                op.keypoints = kpts[j]
            preds.append(op)
        out.append(PredictionResult(object_prediction_list=preds, image=images[i], durations_in_seconds={},
                                    image_size=(w, h)))
    return out
