"""sahi.predict mirror — `get_prediction` / `get_sliced_prediction` with the reference's signatures, defaults and
result objects (docs sahi/predict.py:54-345).

Two execution paths, both on the GPU:
  * fused   — the detection model exposes `supports_batched_slices` (our YOLOv11PoseDetectionModel, or the
              reference's own unmodified utils/yolo_wrapper.py class once its `.model` is an fsd_b200.YOLO): the whole
              slice -> detect -> shift -> merge loop runs as one batched device pipeline (engine.SlicedFaceDetector);
  * generic — any other DetectionModel plug-in (e.g. InsightFace): the plug-in detects slice by slice exactly as in
              the reference, and the shift + merge run through Kernel 3.
"""
from __future__ import annotations

import time
from typing import List, Optional

import numpy as np

from .postprocess import (GreedyNMMPostprocess, LSNMSPostprocess, NMMPostprocess, NMSPostprocess,
                          PostprocessPredictions)
from .prediction import ObjectPrediction, PredictionResult
from .slicing import read_image_as_pil, slice_image

POSTPROCESS_NAME_TO_CLASS = {
    "GREEDYNMM": GreedyNMMPostprocess,
    "NMM": NMMPostprocess,
    "NMS": NMSPostprocess,
    "LSNMS": LSNMSPostprocess,
}

LOW_MODEL_CONFIDENCE = 0.1


def filter_predictions(object_prediction_list, exclude_classes_by_name, exclude_classes_by_id):
    skip_names, skip_ids = exclude_classes_by_name or [], exclude_classes_by_id or []
    return [p for p in object_prediction_list if p.category.name not in skip_names and p.category.id not in skip_ids]


def get_prediction(image, detection_model, shift_amount: list = [0, 0], full_shape=None,
                   postprocess: Optional[PostprocessPredictions] = None, verbose: int = 0,
                   exclude_classes_by_name: Optional[List[str]] = None,
                   exclude_classes_by_id: Optional[List[int]] = None) -> PredictionResult:
    durations = {}
    image_as_pil = read_image_as_pil(image)
    t0 = time.time()
    detection_model.perform_inference(np.ascontiguousarray(image_as_pil))
    durations["prediction"] = time.time() - t0
    if full_shape is None:
        full_shape = [image_as_pil.height, image_as_pil.width]
    t0 = time.time()
    detection_model.convert_original_predictions(shift_amount=shift_amount, full_shape=full_shape)
    preds = filter_predictions(detection_model.object_prediction_list, exclude_classes_by_name, exclude_classes_by_id)
    if postprocess is not None:
        preds = postprocess(preds)
    durations["postprocess"] = time.time() - t0
    if verbose == 1:
        print("Prediction performed in", durations["prediction"], "seconds.")
    return PredictionResult(image=image_as_pil, object_prediction_list=preds, durations_in_seconds=durations)


def _fused_capable(detection_model) -> bool:
    if getattr(detection_model, "supports_batched_slices", False):
        return True
    # the reference's own plug-in class, unmodified, running on top of fsd_b200.YOLO
    from ..yolo import YOLO

    return type(detection_model).__name__ == "YOLOv11PoseDetectionModel" and isinstance(getattr(detection_model, "model", None), YOLO)


def _fused_sliced_prediction(image_as_pil, detection_model, slice_height, slice_width, ov_h, ov_w,
                             perform_standard_pred, postprocess_type, match_metric, match_threshold):
    """Batched device pipeline for one image; fills the plug-in's keypoints_cache like the per-slice loop would."""
    import torch

    from .. import ops

    model = detection_model.model
    dev = ops.resolve_device(getattr(detection_model, "device", "cuda:0"))
    eng = model.engine(dev, getattr(detection_model, "half", True))
    eng.conf, eng.imgsz, eng.truncate = detection_model.confidence_threshold, detection_model.image_size, True
    arr = np.ascontiguousarray(image_as_pil)  # what get_prediction hands to perform_inference (channel order as-is)
    pool = ops.ImagePool.from_numpy([arr], torch.device(dev))
    batch = eng.detect(pool, slice_height, slice_width, ov_h, ov_w, perform_standard_pred, postprocess_type,
                       match_metric, match_threshold, want_stage1=hasattr(detection_model, "keypoints_cache"))
    boxes, scores, kpts, has_k = batch.image(0)
    from .annotation import Category

    face = Category(id=0, name="face")
    ib, sc, hk = boxes.astype(np.int64).tolist(), scores.tolist(), has_k.tolist()
    preds = [ObjectPrediction.from_merged_row(b[0], b[1], b[2], b[3], sc[j], face, kpts[j] if hk[j] else None)
             for j, b in enumerate(ib)]
    if batch.stage1 is not None:  # same side channel the reference fills slice by slice (utils/yolo_wrapper.py:155-162)
        cache = detection_model.keypoints_cache
        for r in batch.stage1["rows"][0]:
            cache[f"{int(r[0])}_{int(r[1])}_{int(r[2])}_{int(r[3])}"] = r[6:21].reshape(5, 3).copy()
    return preds


def get_sliced_prediction(image, detection_model=None, slice_height: Optional[int] = None,
                          slice_width: Optional[int] = None, overlap_height_ratio: float = 0.2,
                          overlap_width_ratio: float = 0.2, perform_standard_pred: bool = True,
                          postprocess_type: str = "GREEDYNMM", postprocess_match_metric: str = "IOS",
                          postprocess_match_threshold: float = 0.5, postprocess_class_agnostic: bool = False,
                          verbose: int = 1, merge_buffer_length: Optional[int] = None,
                          auto_slice_resolution: bool = True, slice_export_prefix: Optional[str] = None,
                          slice_dir: Optional[str] = None, exclude_classes_by_name: Optional[List[str]] = None,
                          exclude_classes_by_id: Optional[List[int]] = None) -> PredictionResult:
    durations = {}
    if postprocess_type not in POSTPROCESS_NAME_TO_CLASS:
        raise ValueError(f"postprocess_type should be one of {list(POSTPROCESS_NAME_TO_CLASS.keys())} but given as {postprocess_type}")
    postprocess = POSTPROCESS_NAME_TO_CLASS[postprocess_type](match_threshold=postprocess_match_threshold,
                                                              match_metric=postprocess_match_metric,
                                                              class_agnostic=postprocess_class_agnostic)
    t_all = time.time()
    image_as_pil = read_image_as_pil(image)
    width, height = image_as_pil.size

    if (_fused_capable(detection_model) and merge_buffer_length is None and slice_height and slice_width
            and postprocess_type != "LSNMS"):
        # ---- fused device path ---------------------------------------------------------------------------
        durations["slice"] = 0.0  # slicing is index arithmetic inside Kernel 1's launch parameters
        preds = _fused_sliced_prediction(image_as_pil, detection_model, slice_height, slice_width,
                                         overlap_height_ratio, overlap_width_ratio, perform_standard_pred,
                                         postprocess_type, postprocess_match_metric, postprocess_match_threshold)
        preds = filter_predictions(preds, exclude_classes_by_name, exclude_classes_by_id)
        if verbose in (1, 2):
            from .. import _cabi

            n = len(_cabi.slice_plan(height, width, slice_height, slice_width, overlap_height_ratio, overlap_width_ratio))
            print(f"Performing prediction on {n} slices.")
        durations["prediction"] = time.time() - t_all
        durations["postprocess"] = 0.0  # merged on the device inside the same pipeline
    else:
        # ---- generic plug-in path: reference control flow, merge in Kernel 3 -----------------------------------
        t0 = time.time()
        sl = slice_image(image=image_as_pil, output_file_name=slice_export_prefix, output_dir=slice_dir,
                         slice_height=slice_height, slice_width=slice_width,
                         overlap_height_ratio=overlap_height_ratio, overlap_width_ratio=overlap_width_ratio,
                         auto_slice_resolution=auto_slice_resolution)
        num_slices = len(sl)
        durations["slice"] = time.time() - t0
        full_shape = [sl.original_image_height, sl.original_image_width]
        post_t = 0.0
        t0 = time.time()
        if verbose in (1, 2):
            print(f"Performing prediction on {num_slices} slices.")
        preds = []
        for i in range(num_slices):
            res = get_prediction(image=sl.images[i], detection_model=detection_model,
                                 shift_amount=sl.starting_pixels[i], full_shape=full_shape,
                                 exclude_classes_by_name=exclude_classes_by_name,
                                 exclude_classes_by_id=exclude_classes_by_id)
            for op in res.object_prediction_list:
                if op:
                    preds.append(op.get_shifted_object_prediction())
            if merge_buffer_length is not None and len(preds) > merge_buffer_length:
                t1 = time.time()
                preds = postprocess(preds)
                post_t += time.time() - t1
        if num_slices > 1 and perform_standard_pred:
            res = get_prediction(image=image_as_pil, detection_model=detection_model, shift_amount=[0, 0],
                                 full_shape=full_shape, postprocess=None,
                                 exclude_classes_by_name=exclude_classes_by_name,
                                 exclude_classes_by_id=exclude_classes_by_id)
            preds.extend(res.object_prediction_list)
        if len(preds) > 1:
            t1 = time.time()
            preds = postprocess(preds)
            post_t += time.time() - t1
        durations["prediction"] = time.time() - t0 - post_t
        durations["postprocess"] = post_t

    if verbose == 2:
        print("Slicing performed in", durations["slice"], "seconds.")
        print("Prediction performed in", durations["prediction"], "seconds.")
        print("Postprocessing performed in", durations["postprocess"], "seconds.")
    return PredictionResult(image=image_as_pil, object_prediction_list=preds, durations_in_seconds=durations,
                            image_size=(width, height))
