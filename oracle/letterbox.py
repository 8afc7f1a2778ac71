"""Oracle for Kernel 1: ultralytics LetterBox + tensor conversion (test infrastructure, see oracle/__init__).

Follows SURVEY.md App. A.3 ([EXT ultralytics] `LetterBox.__call__` and `BasePredictor.preprocess`, reached
from the reference at utils/yolo_wrapper.py:74-80).  Two implementations are kept side by side:
  * `letterbox_cv2`  — the real thing: cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder(114), cv2 is in the image;
  * `resize_linear_u8` — pure-numpy integer restatement of cv2's fixed-point bilinear (and its exact-2x area
    special case); tests pin it bit-exact against cv2 so it can serve where cv2 must not be a dependency.
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is present in this image; the numpy restatement below does not need it
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def letterbox_geometry(src_h: int, src_w: int, imgsz: int = 1024, stride: int = 32):
    """LetterBox(new_shape=imgsz, auto=True, scaleup=True, center=True) geometry.

    Returns dict(new_w,new_h,left,top,right,bottom,out_w,out_h,mode,gain).  `round` is Python's
    round-half-even on a double, exactly as upstream.
    """
    new_shape = (imgsz, imgsz)
    r = min(new_shape[0] / src_h, new_shape[1] / src_w)
    new_w, new_h = int(round(src_w * r)), int(round(src_h * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    dw, dh = float(np.mod(dw, stride)), float(np.mod(dh, stride))
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    if (new_w, new_h) == (src_w, src_h):
        mode = 0
    elif src_w == 2 * new_w and src_h == 2 * new_h:
        mode = 2
    else:
        mode = 1
    out_w, out_h = new_w + left + right, new_h + top + bottom
    gain = min(out_h / src_h, out_w / src_w)  # ultralytics scale_boxes recomputes this from the padded shape
    return dict(new_w=new_w, new_h=new_h, left=left, top=top, right=right, bottom=bottom,
                out_w=out_w, out_h=out_h, mode=mode, gain=gain)


def letterbox_cv2(img: np.ndarray, imgsz: int = 1024, stride: int = 32) -> np.ndarray:
    """HWC uint8 -> letterboxed HWC uint8 using the real cv2 calls of ultralytics.LetterBox."""
    g = letterbox_geometry(img.shape[0], img.shape[1], imgsz, stride)
    if (img.shape[1], img.shape[0]) != (g["new_w"], g["new_h"]):
        img = cv2.resize(img, (g["new_w"], g["new_h"]), interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(img, g["top"], g["bottom"], g["left"], g["right"], cv2.BORDER_CONSTANT,
                              value=(114, 114, 114))


def _axis_tables(src: int, dst: int, is_x: bool):
    """cv2 resize.cpp coefficient tables for the uint8 INTER_LINEAR path (float32 arithmetic like cv2)."""
    inv_scale = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if is_x:
        lo = s < 0
        s[lo] = 0
        f[lo] = 0
        hi = s >= src - 1
        s[hi] = src - 1
        f[hi] = 0
        s0, s1 = s, np.minimum(s + 1, src - 1)
    else:
        s0, s1 = np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1)
    w0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    w1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return s0, s1, w0, w1


def resize_linear_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """Integer restatement of cv2.resize(img, (new_w,new_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC."""
    h, w = img.shape[:2]
    if (w, h) == (new_w, new_h):
        return img.copy()
    src = img.astype(np.int64)
    if w == 2 * new_w and h == 2 * new_h:  # cv2 switches INTER_LINEAR to the INTER_AREA 2x2 fast path
        a = src[0::2, 0::2] + src[0::2, 1::2] + src[1::2, 0::2] + src[1::2, 1::2]
        return ((a + 2) >> 2).astype(np.uint8)
    sx0, sx1, ax0, ax1 = _axis_tables(w, new_w, True)
    sy0, sy1, by0, by1 = _axis_tables(h, new_h, False)
    rows = src[:, sx0] * ax0[None, :, None] + src[:, sx1] * ax1[None, :, None]  # [h,new_w,3] int
    r0 = rows[sy0] >> 4
    r1 = rows[sy1] >> 4
    out = (((by0[:, None, None] * r0) >> 16) + ((by1[:, None, None] * r1) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_numpy(img: np.ndarray, imgsz: int = 1024, stride: int = 32) -> np.ndarray:
    g = letterbox_geometry(img.shape[0], img.shape[1], imgsz, stride)
    res = resize_linear_u8(img, g["new_w"], g["new_h"])
    out = np.full((g["out_h"], g["out_w"], 3), 114, dtype=np.uint8)
    out[g["top"]:g["top"] + g["new_h"], g["left"]:g["left"] + g["new_w"]] = res
    return out


def preprocess(img_hwc_u8: np.ndarray, imgsz: int = 1024, stride: int = 32, half: bool = False,
               use_cv2: bool = True):
    """ultralytics BasePredictor.preprocess for one ndarray source: letterbox, [..., ::-1], HWC->CHW, /255.

    Returns a torch tensor [1,3,H,W] (fp32, or fp16 when half=True: `im.half(); im /= 255`).
    """
    import torch

    lb = letterbox_cv2(img_hwc_u8, imgsz, stride) if use_cv2 else letterbox_numpy(img_hwc_u8, imgsz, stride)
    im = np.ascontiguousarray(lb[None][..., ::-1].transpose(0, 3, 1, 2))
    t = torch.from_numpy(im)
    t = t.half() if half else t.float()
    t /= 255
    return t
