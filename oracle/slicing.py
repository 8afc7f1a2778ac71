"""Oracle for the slice planner (test infrastructure, see oracle/__init__).

Restates [EXT sahi==0.11.34] `sahi.slicing.get_slice_bboxes` / `slice_image` and `sahi.utils.cv.read_image_as_pil`
as recorded in SURVEY.md App. A.1; called by the reference at docs sahi/predict.py:229-241 and
scripts/debug_slicing.py:74-90.  Parity unpinned (upstream source absent); pinned only by the hand-derived
known-answer table of SURVEY App. B (tests/test_oracle_slicing.py).
"""
from __future__ import annotations

import numpy as np
from PIL import Image, ImageOps


def get_slice_bboxes(image_height, image_width, slice_height=None, slice_width=None, auto_slice_resolution=True,
                     overlap_height_ratio=0.2, overlap_width_ratio=0.2):
    """Row-major [x0,y0,x1,y1] windows; overlap px = int(ratio*size); border windows are shifted back inside."""
    if not (slice_height and slice_width):
        raise ValueError("auto slice resolution is out of scope: no reference caller omits the slice size")
    y_overlap = int(overlap_height_ratio * slice_height)
    x_overlap = int(overlap_width_ratio * slice_width)
    out = []
    y_min = y_max = 0
    while y_max < image_height:
        x_min = x_max = 0
        y_max = y_min + slice_height
        while x_max < image_width:
            x_max = x_min + slice_width
            if y_max > image_height or x_max > image_width:
                xe, ye = min(image_width, x_max), min(image_height, y_max)
                out.append([max(0, xe - slice_width), max(0, ye - slice_height), xe, ye])
            else:
                out.append([x_min, y_min, x_max, y_max])
            x_min = x_max - x_overlap
        y_min = y_max - y_overlap
    return out


def read_image_as_pil(image):
    """str -> RGB PIL (EXIF-transposed); ndarray -> Image.fromarray as-is (no BGR fix); PIL -> unchanged."""
    if isinstance(image, Image.Image):
        return image
    if isinstance(image, str):
        return ImageOps.exif_transpose(Image.open(image).convert("RGB"))
    if isinstance(image, np.ndarray):
        if image.shape[0] < 5:  # CHW given: upstream flips to HWC via [:, :, ::-1]
            image = image[:, :, ::-1]
        return Image.fromarray(image)
    raise TypeError("image must be a path, PIL image or numpy array")


class SliceImageResult:
    def __init__(self, original_image_size):
        self.original_image_height = int(original_image_size[0])
        self.original_image_width = int(original_image_size[1])
        self.images = []
        self.starting_pixels = []

    def __len__(self):
        return len(self.images)


def slice_image(image, slice_height=None, slice_width=None, overlap_height_ratio=0.2, overlap_width_ratio=0.2,
                output_dir=None, output_file_name=None, auto_slice_resolution=True, **_):
    pil = read_image_as_pil(image)
    w, h = pil.size
    arr = np.asarray(pil)
    res = SliceImageResult([h, w])
    for x0, y0, x1, y1 in get_slice_bboxes(h, w, slice_height, slice_width, auto_slice_resolution,
                                           overlap_height_ratio, overlap_width_ratio):
        res.images.append(arr[y0:y1, x0:x1])
        res.starting_pixels.append([x0, y0])
    return res
