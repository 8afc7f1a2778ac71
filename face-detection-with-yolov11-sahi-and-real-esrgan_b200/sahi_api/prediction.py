"""sahi.prediction mirror (reference: docs sahi/prediction.py:13-243): PredictionScore / ObjectPrediction /
PredictionResult.  Host-side boundary objects; PredictionResult decodes its image lazily (the reference decodes a
path-given image a second time in the constructor only to learn its size)."""
from __future__ import annotations

import copy
from typing import Any, Dict, List, Optional

import numpy as np

from .annotation import BoundingBox, Category, ObjectAnnotation
from .slicing import read_image_as_pil


class PredictionScore:
    def __init__(self, value):
        self.value = copy.deepcopy(value).tolist() if type(value).__module__ == "numpy" else value

    def is_greater_than_threshold(self, threshold):
        return self.value > threshold

    def __eq__(self, threshold):
        return self.value == threshold

    def __gt__(self, threshold):
        return self.value > threshold

    def __lt__(self, threshold):
        return self.value < threshold

    def __repr__(self):
        return f"PredictionScore: <value: {self.value}>"


class ObjectPrediction(ObjectAnnotation):
    def __init__(self, bbox: Optional[List[int]] = None, category_id: Optional[int] = None,
                 category_name: Optional[str] = None, segmentation=None, score: float = 0.0,
                 shift_amount: Optional[List[int]] = [0, 0], full_shape: Optional[List[int]] = None):
        self.score = PredictionScore(score)
        super().__init__(bbox=bbox, category_id=category_id, segmentation=segmentation, category_name=category_name,
                         shift_amount=shift_amount, full_shape=full_shape)

    @classmethod
    def from_merged_row(cls, x1: int, y1: int, x2: int, y2: int, score: float, category: Category, keypoints=None):
        """Fast construction for rows coming back from the device pipeline: the values already satisfy every rule the
        regular constructor enforces (non-negative ints in full-image coordinates, shift (0, 0), no full_shape clamp),
        so the per-object validation is skipped — a batch returns thousands of these."""
        self = object.__new__(cls)
        box = object.__new__(BoundingBox)
        object.__setattr__(box, "box", [x1, y1, x2, y2])
        object.__setattr__(box, "shift_amount", (0, 0))
        score_obj = object.__new__(PredictionScore)
        score_obj.value = score
        self.score, self.mask, self.bbox, self.category, self.merged = score_obj, None, box, category, None
        if keypoints is not None:
            self.keypoints = keypoints
        return self

    def get_shifted_object_prediction(self):
        """Copy mapped into full-image coordinates (shift applied, shift_amount reset, no upper clamp)."""
        return ObjectPrediction(bbox=self.bbox.get_shifted_box().to_xyxy(), category_id=self.category.id,
                                score=self.score.value, segmentation=None, category_name=self.category.name,
                                shift_amount=[0, 0], full_shape=None)

    def to_coco_prediction(self, image_id=None):
        return _CocoPrediction(self.bbox.to_xywh(), self.category.id, self.category.name, self.score.value, image_id)

    def __repr__(self):
        return f"ObjectPrediction<\n    bbox: {self.bbox},\n    mask: {self.mask},\n    score: {self.score},\n    category: {self.category}>"


class _CocoPrediction:
    def __init__(self, bbox, category_id, category_name, score, image_id):
        self.json = {"image_id": image_id, "bbox": bbox, "score": score, "category_id": category_id,
                     "category_name": category_name, "segmentation": [], "iscrowd": 0,
                     "area": bbox[2] * bbox[3]}


class PredictionResult:
    def __init__(self, object_prediction_list: List[ObjectPrediction], image, durations_in_seconds: Dict[str, Any] = dict(),
                 image_size=None):
        self._image_src = image
        self._image = None
        self._size = image_size  # (width, height) when the caller already knows it
        self.object_prediction_list: List[ObjectPrediction] = object_prediction_list
        self.durations_in_seconds = durations_in_seconds

    @property
    def image(self):
        if self._image is None:
            self._image = read_image_as_pil(self._image_src)
        return self._image

    @image.setter
    def image(self, value):
        self._image = value

    @property
    def image_width(self):
        return self._size[0] if self._size else self.image.size[0]

    @property
    def image_height(self):
        return self._size[1] if self._size else self.image.size[1]

    def export_visuals(self, export_dir: str, text_size=None, rect_th=None, hide_labels=False, hide_conf=False,
                       file_name: str = "prediction_visual"):
        import os

        import cv2

        os.makedirs(export_dir, exist_ok=True)
        canvas = np.ascontiguousarray(self.image).copy()
        th = rect_th or max(round(sum(canvas.shape[:2]) / 2 * 0.003), 2)
        for op in self.object_prediction_list:
            x1, y1, x2, y2 = (int(v) for v in op.bbox.to_xyxy())
            cv2.rectangle(canvas, (x1, y1), (x2, y2), (0, 200, 0), th)
            if not hide_labels:
                label = op.category.name if hide_conf else f"{op.category.name} {op.score.value:.2f}"
                cv2.putText(canvas, label, (x1, max(0, y1 - 3)), 0, text_size or th / 3, (0, 200, 0), max(th - 1, 1))
        cv2.imwrite(os.path.join(export_dir, file_name + ".png"), cv2.cvtColor(canvas, cv2.COLOR_RGB2BGR))

    def to_coco_annotations(self):
        return [op.to_coco_prediction().json for op in self.object_prediction_list]

    def to_coco_predictions(self, image_id: Optional[int] = None):
        return [op.to_coco_prediction(image_id=image_id).json for op in self.object_prediction_list]
