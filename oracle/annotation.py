"""Oracle data classes (test infrastructure, see oracle/__init__).

`BoundingBox`, `Category`, `ObjectAnnotation` restate [EXT sahi==0.11.34] sahi.annotation;
`PredictionScore`, `ObjectPrediction`, `PredictionResult` restate the reference's vendored
docs sahi/prediction.py:13-176 (checked against it by tests/golden).
"""
from __future__ import annotations

import copy

import numpy as np

from .slicing import read_image_as_pil


class BoundingBox:
    """xyxy box + the (shift_x, shift_y) that maps it from slice to full-image coordinates."""

    def __init__(self, box, shift_amount=(0, 0)):
        if box[0] < 0 or box[1] < 0 or box[2] < 0 or box[3] < 0:
            raise Exception("Box coords [minx, miny, maxx, maxy] cannot be negative")
        self.minx, self.miny, self.maxx, self.maxy = box[0], box[1], box[2], box[3]
        self.box = [self.minx, self.miny, self.maxx, self.maxy]
        self.shift_amount = tuple(shift_amount)
        self.shift_x, self.shift_y = shift_amount[0], shift_amount[1]

    @property
    def area(self):
        return (self.maxx - self.minx) * (self.maxy - self.miny)

    def to_xyxy(self):
        return [self.minx, self.miny, self.maxx, self.maxy]

    to_voc_bbox = to_xyxy

    def to_xywh(self):
        return [self.minx, self.miny, self.maxx - self.minx, self.maxy - self.miny]

    to_coco_bbox = to_xywh

    def get_shifted_box(self):
        return BoundingBox([self.minx + self.shift_x, self.miny + self.shift_y, self.maxx + self.shift_x,
                            self.maxy + self.shift_y], shift_amount=(0, 0))

    def get_expanded_box(self, ratio=0.1, max_x=None, max_y=None):
        w, h = self.maxx - self.minx, self.maxy - self.miny
        dx, dy = int(w * ratio), int(h * ratio)
        x1 = min(max_x, self.maxx + dx) if max_x else self.maxx + dx
        y1 = min(max_y, self.maxy + dy) if max_y else self.maxy + dy
        return BoundingBox([max(0, self.minx - dx), max(0, self.miny - dy), x1, y1], shift_amount=self.shift_amount)

    def __repr__(self):
        return f"BoundingBox: <{(self.minx, self.miny, self.maxx, self.maxy)}, w: {self.maxx - self.minx}, h: {self.maxy - self.miny}>"


class Category:
    def __init__(self, id=None, name=None):
        self.id, self.name = id, name

    def __repr__(self):
        return f"Category: <id: {self.id}, name: {self.name}>"


class ObjectAnnotation:
    """Clamp rule of sahi ObjectAnnotation.__init__: min >= 0, max <= full_shape (slice-local vs full shape)."""

    def __init__(self, bbox=None, segmentation=None, category_id=None, category_name=None, shift_amount=[0, 0],
                 full_shape=None):
        if segmentation is not None:
            raise NotImplementedError("masks are out of scope (has_mask is False for every reference plugin)")
        if bbox is None:
            raise ValueError("you must provide a bbox")
        if type(bbox).__module__ == "numpy":
            bbox = copy.deepcopy(bbox).tolist()
        xmin, ymin = max(bbox[0], 0), max(bbox[1], 0)
        if full_shape:
            xmax, ymax = min(bbox[2], full_shape[1]), min(bbox[3], full_shape[0])
        else:
            xmax, ymax = bbox[2], bbox[3]
        self.mask = None
        self.bbox = BoundingBox([xmin, ymin, xmax, ymax], shift_amount)
        self.category = Category(id=category_id, name=category_name if category_name else str(category_id))
        self.merged = None


class PredictionScore:  # docs sahi/prediction.py:13-41
    def __init__(self, value):
        if type(value).__module__ == "numpy":
            value = copy.deepcopy(value).tolist()
        self.value = value

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


class ObjectPrediction(ObjectAnnotation):  # docs sahi/prediction.py:44-163
    def __init__(self, bbox=None, category_id=None, category_name=None, segmentation=None, score=0.0,
                 shift_amount=[0, 0], full_shape=None):
        self.score = PredictionScore(score)
        super().__init__(bbox=bbox, category_id=category_id, segmentation=segmentation,
                         category_name=category_name, shift_amount=shift_amount, full_shape=full_shape)

    def get_shifted_object_prediction(self):  # :94-120 (mask branch out of scope)
        return ObjectPrediction(bbox=self.bbox.get_shifted_box().to_xyxy(), category_id=self.category.id,
                                score=self.score.value, segmentation=None, category_name=self.category.name,
                                shift_amount=[0, 0], full_shape=None)

    def __repr__(self):
        return f"ObjectPrediction<bbox: {self.bbox}, score: {self.score}, category: {self.category}>"


class PredictionResult:  # docs sahi/prediction.py:166-176
    def __init__(self, object_prediction_list, image, durations_in_seconds=dict()):
        self.image = read_image_as_pil(image)
        self.image_width, self.image_height = self.image.size
        self.object_prediction_list = object_prediction_list
        self.durations_in_seconds = durations_in_seconds
