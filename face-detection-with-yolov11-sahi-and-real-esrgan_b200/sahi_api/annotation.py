"""sahi.annotation mirror: BoundingBox / Category / ObjectAnnotation (clamp rules of SURVEY App. A.2.1).

Host-side value objects only — they are the boundary objects reference callers read (`det.bbox.to_xyxy()`,
`det.category.name`, utils/visualization.py:107-133); no arithmetic of the hot path lives here."""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple


@dataclass(frozen=True)
class BoundingBox:
    box: Sequence[float]
    shift_amount: Tuple[int, int] = (0, 0)

    def __post_init__(self):
        if len(self.box) != 4 or any(c < 0 for c in self.box):
            raise Exception("box must be 4 non-negative floats: [minx, miny, maxx, maxy]")
        if len(self.shift_amount) != 2:
            raise ValueError("shift_amount must be 2 integers: [shift_x, shift_y]")
        object.__setattr__(self, "box", [self.box[0], self.box[1], self.box[2], self.box[3]])
        object.__setattr__(self, "shift_amount", (self.shift_amount[0], self.shift_amount[1]))

    minx = property(lambda self: self.box[0])
    miny = property(lambda self: self.box[1])
    maxx = property(lambda self: self.box[2])
    maxy = property(lambda self: self.box[3])
    shift_x = property(lambda self: self.shift_amount[0])
    shift_y = property(lambda self: self.shift_amount[1])

    @property
    def area(self):
        return (self.maxx - self.minx) * (self.maxy - self.miny)

    def get_expanded_box(self, ratio: float = 0.1, max_x: Optional[int] = None, max_y: Optional[int] = None):
        dx, dy = int((self.maxx - self.minx) * ratio), int((self.maxy - self.miny) * ratio)
        hi_x = min(max_x, self.maxx + dx) if max_x else self.maxx + dx
        hi_y = min(max_y, self.maxy + dy) if max_y else self.maxy + dy
        return BoundingBox([max(0, self.minx - dx), max(0, self.miny - dy), hi_x, hi_y], self.shift_amount)

    def to_xywh(self):
        return [self.minx, self.miny, self.maxx - self.minx, self.maxy - self.miny]

    def to_coco_bbox(self):
        return self.to_xywh()

    def to_xyxy(self):
        return [self.minx, self.miny, self.maxx, self.maxy]

    def to_voc_bbox(self):
        return self.to_xyxy()

    def get_shifted_box(self):
        sx, sy = self.shift_amount
        return BoundingBox([self.minx + sx, self.miny + sy, self.maxx + sx, self.maxy + sy], (0, 0))

    def __repr__(self):
        return (f"BoundingBox: <{(self.minx, self.miny, self.maxx, self.maxy)}, "
                f"w: {self.maxx - self.minx}, h: {self.maxy - self.miny}>")


@dataclass(frozen=True)
class Category:
    id: Optional[int] = None
    name: Optional[str] = None

    def __repr__(self):
        return f"Category: <id: {self.id}, name: {self.name}>"


class ObjectAnnotation:
    """Box + category.  The upper clamp compares slice-local coordinates with the FULL image shape, as upstream."""

    def __init__(self, bbox: Optional[List[int]] = None, segmentation=None, category_id: Optional[int] = None,
                 category_name: Optional[str] = None, shift_amount: Optional[List[int]] = [0, 0],
                 full_shape: Optional[List[int]] = None):
        if not isinstance(category_id, int):
            raise ValueError("category_id must be an integer")
        if segmentation is not None:
            raise NotImplementedError("segmentation masks are outside this path (has_mask is False for every plugin)")
        if bbox is None:
            raise ValueError("you must provide a bbox")
        if type(bbox).__module__ == "numpy":
            bbox = copy.deepcopy(bbox).tolist()
        lo_x, lo_y = max(bbox[0], 0), max(bbox[1], 0)
        hi_x, hi_y = (min(bbox[2], full_shape[1]), min(bbox[3], full_shape[0])) if full_shape else (bbox[2], bbox[3])
        self.mask = None
        self.bbox = BoundingBox([lo_x, lo_y, hi_x, hi_y], shift_amount)
        self.category = Category(id=category_id, name=category_name if category_name else str(category_id))
        self.merged = None

    def deepcopy(self):
        return copy.deepcopy(self)

    def __repr__(self):
        return f"ObjectAnnotation<bbox: {self.bbox}, mask: {self.mask}, category: {self.category}>"
