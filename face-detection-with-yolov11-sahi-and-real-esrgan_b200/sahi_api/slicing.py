"""sahi.slicing / sahi.utils.cv mirror: the slice grid comes from the C library's integer planner
(`fsd_slice_plan`); `slice_image` keeps the reference-visible result object (.images, .starting_pixels, ...)."""
from __future__ import annotations

import numpy as np
from PIL import Image, ImageOps

from .. import _cabi


def get_slice_bboxes(image_height: int, image_width: int, slice_height=None, slice_width=None,
                     auto_slice_resolution: bool = True, overlap_height_ratio: float = 0.2,
                     overlap_width_ratio: float = 0.2):
    if not (slice_height and slice_width):
        if auto_slice_resolution:
            raise NotImplementedError("automatic slice resolution is not on this path: pass slice_height/slice_width "
                                      "(every reference caller does)")
        raise ValueError("Compute type is not auto and slice width and height are not provided.")
    return _cabi.slice_plan(int(image_height), int(image_width), int(slice_height), int(slice_width),
                            float(overlap_height_ratio), float(overlap_width_ratio))


def read_image_as_pil(image, exif_fix: bool = True):
    if isinstance(image, Image.Image):
        return image
    if isinstance(image, str):
        pil = Image.open(image).convert("RGB")
        return ImageOps.exif_transpose(pil) if exif_fix else pil
    if isinstance(image, (bytes, bytearray)):  # (f4) an encoded image kept behind a result of the nvJPEG batch path: decoded lazily, on request
        import io

        pil = Image.open(io.BytesIO(image)).convert("RGB")
        return ImageOps.exif_transpose(pil) if exif_fix else pil
    if isinstance(image, np.ndarray):
        if image.shape[0] < 5:  # upstream treats a leading dim < 5 as CHW
            image = image[:, :, ::-1]
        return Image.fromarray(image)
    raise TypeError("read image with 'pillow' using 'Image.open()'")


class SliceImageResult:
    def __init__(self, original_image_size, image_dir=None):
        self.original_image_height = int(original_image_size[0])
        self.original_image_width = int(original_image_size[1])
        self.image_dir = image_dir
        self.images, self.starting_pixels = [], []

    @property
    def filenames(self):
        return [None] * len(self.images)

    def __len__(self):
        return len(self.images)


def slice_image(image, coco_annotation_list=None, output_file_name=None, output_dir=None, slice_height=None,
                slice_width=None, overlap_height_ratio: float = 0.2, overlap_width_ratio: float = 0.2,
                auto_slice_resolution: bool = True, min_area_ratio: float = 0.1, out_ext=None, verbose: bool = False):
    pil = read_image_as_pil(image)
    w, h = pil.size
    arr = np.asarray(pil)
    out = SliceImageResult([h, w], output_dir)
    for x0, y0, x1, y1 in get_slice_bboxes(h, w, slice_height, slice_width, auto_slice_resolution,
                                           overlap_height_ratio, overlap_width_ratio):
        out.images.append(arr[y0:y1, x0:x1])
        out.starting_pixels.append([x0, y0])
    return out
