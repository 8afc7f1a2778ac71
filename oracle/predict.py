"""Oracle orchestration (test infrastructure, see oracle/__init__).

Restates the reference's vendored SAHI driver, docs sahi/predict.py:54-345 (`filter_predictions`,
`get_prediction`, `get_sliced_prediction`) and the plugin base docs sahi/base.py:12-196, with the reference's
exact control flow: sequential batch-1 slices, per-box Python objects, CPU merge.  This file is pinned: the
reference's own predict.py is imported unmodified by tests/golden/make_golden.py and must give the same result.
"""
from __future__ import annotations

import time

import numpy as np

from .annotation import Category, ObjectPrediction, PredictionResult
from .postprocess import POSTPROCESS_NAME_TO_CLASS
from .slicing import read_image_as_pil, slice_image


class DetectionModel:
    """Plugin protocol of docs sahi/base.py:12-196 (device selection reduced to storing the string)."""

    def __init__(self, model_path=None, model=None, config_path=None, device=None, mask_threshold=0.5,
                 confidence_threshold=0.3, category_mapping=None, category_remapping=None, load_at_init=True,
                 image_size=None):
        self.model_path, self.config_path, self.model = model_path, config_path, None
        self.mask_threshold, self.confidence_threshold = mask_threshold, confidence_threshold
        self.category_mapping, self.category_remapping = category_mapping, category_remapping
        self.image_size = image_size
        self._original_predictions = None
        self._object_prediction_list_per_image = None
        self.set_device(device)
        if load_at_init:
            if model:
                self.set_model(model)
            else:
                self.load_model()

    def load_model(self):
        raise NotImplementedError()

    def set_model(self, model, **kwargs):
        raise NotImplementedError()

    def set_device(self, device=None):
        self.device = device if device is not None else "cpu"

    def unload_model(self):
        self.model = None

    def perform_inference(self, image):
        raise NotImplementedError()

    def _create_object_prediction_list_from_original_predictions(self, shift_amount_list=[[0, 0]],
                                                                 full_shape_list=None):
        raise NotImplementedError()

    def _apply_category_remapping(self):
        if self.category_remapping is None:
            raise ValueError("self.category_remapping cannot be None")
        for lst in self._object_prediction_list_per_image:
            for op in lst:
                op.category = Category(id=self.category_remapping[str(op.category.id)], name=op.category.name)

    def convert_original_predictions(self, shift_amount=[[0, 0]], full_shape=None):
        self._create_object_prediction_list_from_original_predictions(shift_amount_list=shift_amount,
                                                                      full_shape_list=full_shape)
        if self.category_remapping:
            self._apply_category_remapping()

    @property
    def object_prediction_list(self):
        if not self._object_prediction_list_per_image:
            return []
        return self._object_prediction_list_per_image[0]

    @property
    def object_prediction_list_per_image(self):
        return self._object_prediction_list_per_image or []

    @property
    def original_predictions(self):
        return self._original_predictions


def filter_predictions(object_prediction_list, exclude_classes_by_name, exclude_classes_by_id):
    return [p for p in object_prediction_list
            if p.category.name not in (exclude_classes_by_name or []) and p.category.id not in (exclude_classes_by_id or [])]


def get_prediction(image, detection_model, shift_amount=[0, 0], full_shape=None, postprocess=None, verbose=0,
                   exclude_classes_by_name=None, exclude_classes_by_id=None):
    durations = {}
    pil = read_image_as_pil(image)
    t0 = time.time()
    detection_model.perform_inference(np.ascontiguousarray(pil))
    durations["prediction"] = time.time() - t0
    if full_shape is None:
        full_shape = [pil.height, pil.width]
    t0 = time.time()
    detection_model.convert_original_predictions(shift_amount=shift_amount, full_shape=full_shape)
    preds = filter_predictions(detection_model.object_prediction_list, exclude_classes_by_name, exclude_classes_by_id)
    if postprocess is not None:
        preds = postprocess(preds)
    durations["postprocess"] = time.time() - t0
    if verbose == 1:
        print("Prediction performed in", durations["prediction"], "seconds.")
    return PredictionResult(image=image, object_prediction_list=preds, durations_in_seconds=durations)


def get_sliced_prediction(image, detection_model=None, slice_height=None, slice_width=None,
                          overlap_height_ratio=0.2, overlap_width_ratio=0.2, perform_standard_pred=True,
                          postprocess_type="GREEDYNMM", postprocess_match_metric="IOS",
                          postprocess_match_threshold=0.5, postprocess_class_agnostic=False, verbose=1,
                          merge_buffer_length=None, auto_slice_resolution=True, slice_export_prefix=None,
                          slice_dir=None, exclude_classes_by_name=None, exclude_classes_by_id=None):
    durations = {}
    t0 = time.time()
    sl = slice_image(image=image, output_file_name=slice_export_prefix, output_dir=slice_dir,
                     slice_height=slice_height, slice_width=slice_width,
                     overlap_height_ratio=overlap_height_ratio, overlap_width_ratio=overlap_width_ratio,
                     auto_slice_resolution=auto_slice_resolution)
    num_slices = len(sl)
    durations["slice"] = time.time() - t0
    if postprocess_type not in POSTPROCESS_NAME_TO_CLASS:
        raise ValueError(f"postprocess_type should be one of {list(POSTPROCESS_NAME_TO_CLASS)} but given as {postprocess_type}")
    postprocess = POSTPROCESS_NAME_TO_CLASS[postprocess_type](match_threshold=postprocess_match_threshold,
                                                              match_metric=postprocess_match_metric,
                                                              class_agnostic=postprocess_class_agnostic)
    full_shape = [sl.original_image_height, sl.original_image_width]
    post_t = 0.0
    t0 = time.time()
    if verbose in (1, 2):
        print(f"Performing prediction on {num_slices} slices.")
    preds = []
    for i in range(num_slices):  # reference: num_batch = 1, strictly sequential
        res = get_prediction(image=sl.images[i], detection_model=detection_model,
                             shift_amount=sl.starting_pixels[i], full_shape=full_shape,
                             exclude_classes_by_name=exclude_classes_by_name, exclude_classes_by_id=exclude_classes_by_id)
        for op in res.object_prediction_list:
            if op:
                preds.append(op.get_shifted_object_prediction())
        if merge_buffer_length is not None and len(preds) > merge_buffer_length:
            t1 = time.time()
            preds = postprocess(preds)
            post_t += time.time() - t1
    if num_slices > 1 and perform_standard_pred:
        res = get_prediction(image=image, detection_model=detection_model, shift_amount=[0, 0], full_shape=full_shape,
                             postprocess=None, exclude_classes_by_name=exclude_classes_by_name,
                             exclude_classes_by_id=exclude_classes_by_id)
        preds.extend(res.object_prediction_list)
    if len(preds) > 1:
        t1 = time.time()
        preds = postprocess(preds)
        post_t += time.time() - t1
    total = time.time() - t0
    durations["prediction"] = total - post_t
    durations["postprocess"] = post_t
    if verbose == 2:
        print("Slicing performed in", durations["slice"], "seconds.")
        print("Prediction performed in", durations["prediction"], "seconds.")
        print("Postprocessing performed in", durations["postprocess"], "seconds.")
    return PredictionResult(image=image, object_prediction_list=preds, durations_in_seconds=durations)
