"""(f3) Pipeline glue of the reference's v1 / v2 / evaluator scripts as host helpers + device-resident pipelines, so the
enhancement-first and detection-first flows never round-trip through temporary JPEG/PNG files
(reference: pipeline_v4_yolo/app_yolo_full.py:103-104, pipeline_v2_enhancement_first/app_v2.py:105-106,
pipeline_v4_yolo/1_Inference.py:328-330 all write the intermediate image to disk and hand SAHI the path)."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np
import torch

from . import ops
from .sahi_api.annotation import BoundingBox
from .sahi_api.prediction import PredictionResult


def choose_slice_params(img_w: int, img_h: int, prefer: str = "auto"):
    """pipeline_v2_enhancement_first/app_v2.py:19-45 — 3x3 grid below 3000 px on the long side, else 4x4; slice sizes
    rounded up to a multiple of 64 and capped at the image; overlap 0.2.  Returns (slice_h, slice_w, ov_h, ov_w)."""
    long_side = max(img_w, img_h)
    cols = rows = 3 if (prefer == "3x3" or (prefer == "auto" and long_side < 3000)) else 4
    round64 = lambda x: int(math.ceil(x / 64) * 64)  # noqa: E731
    slice_w = min(round64(math.ceil(img_w / cols)), img_w)
    slice_h = min(round64(math.ceil(img_h / rows)), img_h)
    return slice_h, slice_w, 0.2, 0.2


def adaptive_slice_size(w: int, h: int) -> int:
    """eval/eval_official_widerface.py:160-164 — 512 / 416 / 320 by the longer image side."""
    m = max(w, h)
    return 512 if m > 2500 else (416 if m > 1500 else 320)


def detection_first_slice_size(dim: int, base: int = 512) -> int:
    """pipeline_v1_detection_first/app_v1.py:44-51 — half the dimension for small images, else 512 (never below 1)."""
    return max(dim // 2 if dim < base * 1.5 else base, 1)


def crop_rectangles(boxes_xyxy, image_w: int, image_h: int) -> List[Tuple[int, int, int, int]]:
    """utils/visualization.py:206-221 (save_face_crops): int() of the box, clamped to the image; empty crops are skipped."""
    out = []
    for b in boxes_xyxy:
        x1, y1, x2, y2 = (int(c) for c in b)
        x1, y1, x2, y2 = max(0, x1), max(0, y1), min(image_w, x2), min(image_h, y2)
        if x2 > x1 and y2 > y1:
            out.append((x1, y1, x2, y2))
    return out


def rescale_boxes_(prediction_list, scale: float):
    """app_v2.py:131-144: boxes detected on the enhanced image -> original coordinates (float division, new BoundingBox)."""
    for pred in prediction_list:
        x1, y1, x2, y2 = pred.bbox.to_xyxy()
        pred.bbox = BoundingBox([x1 / scale, y1 / scale, x2 / scale, y2 / scale])
    return prediction_list


@torch.no_grad()
def enhance_then_detect(image_bgr: np.ndarray, enhancer, detection_model, prefer: str = "auto", slice_params=None,
                        postprocess_type="GREEDYNMM", match_metric="IOS", match_threshold=0.5, rescale=True):
    """Config 5 / pipeline v2: Real-ESRGAN up-scale of the whole image, then sliced detection on the enhanced image, all
    on the device: Kernel 4 crop -> RRDBNet -> Kernel 4 stitch -> (the stitched uint8 image stays in HBM) -> image pool ->
    Kernel 1 ... -> merged boxes.  `enhancer` is an fsd_b200 RealESRGANer / FaceEnhancer.upsampler."""
    up = getattr(enhancer, "upsampler", enhancer)
    dev = up.device
    big = up.enhance_device(torch.from_numpy(np.ascontiguousarray(image_bgr)).to(dev))  # [H*s, W*s, 3] u8 BGR
    H, W = int(big.shape[0]), int(big.shape[1])
    sh, sw, ovh, ovw = slice_params if slice_params is not None else choose_slice_params(W, H, prefer)
    pool = ops.ImagePool(1, H, W, dev)
    # the evaluator hands SAHI the BGR array as it is (eval_official_widerface.py:184-201); the CLIs re-read the saved file
    # (true RGB).  We follow the evaluator: no channel swap.
    pool.buf[0, :, : W * 3].copy_(big.reshape(H, W * 3))
    eng = detection_model.engine()
    eng.truncate = True
    batch = eng.detect(pool, sh, sw, ovh, ovw, True, postprocess_type, match_metric, match_threshold)
    boxes, scores, kpts, has_k = batch.image(0)
    from .api import _FACE
    from .sahi_api.prediction import ObjectPrediction

    preds = [ObjectPrediction.from_merged_row(int(b[0]), int(b[1]), int(b[2]), int(b[3]), float(s), _FACE, k if hk else None)
             for b, s, k, hk in zip(boxes, scores, kpts, has_k)]
    if rescale:
        rescale_boxes_(preds, float(up.scale))
    return PredictionResult(object_prediction_list=preds, image=image_bgr, durations_in_seconds={},
                            image_size=(image_bgr.shape[1], image_bgr.shape[0])), big


@torch.no_grad()
def detect_then_enhance(image: np.ndarray, detection_model, enhancer, slice_params=None, postprocess_type="GREEDYNMM",
                        match_metric="IOS", match_threshold=0.5):
    """Config 4 / pipeline v1: sliced detection, then Real-ESRGAN of every detected face crop (crop rules of
    save_face_crops), without writing crops to disk.  Returns (PredictionResult, [enhanced crop arrays])."""
    from .sahi_api.predict import get_sliced_prediction

    H, W = image.shape[:2]
    sh, sw = slice_params if slice_params is not None else (detection_first_slice_size(H), detection_first_slice_size(W))
    res = get_sliced_prediction(image, detection_model, slice_height=sh, slice_width=sw, overlap_height_ratio=0.2,
                                overlap_width_ratio=0.2, postprocess_type=postprocess_type,
                                postprocess_match_metric=match_metric, postprocess_match_threshold=match_threshold, verbose=0)
    up = getattr(enhancer, "upsampler", enhancer)
    dev_img = torch.from_numpy(np.ascontiguousarray(image)).to(up.device)
    crops = []
    for (x1, y1, x2, y2) in crop_rectangles([p.bbox.to_xyxy() for p in res.object_prediction_list], W, H):
        if y2 - y1 < 4 or x2 - x1 < 4:  # FaceEnhancer.enhance_image skips images below 4 px (utils/enhancer.py:205-208)
            crops.append(image[y1:y2, x1:x2].copy())
            continue
        crop = dev_img[y1:y2, x1:x2].contiguous()
        crops.append(up.enhance_device(crop).cpu().numpy())
    return res, crops
