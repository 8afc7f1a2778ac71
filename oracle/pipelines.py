"""Oracle for the pipeline glue (test infrastructure): restates pipeline_v2_enhancement_first/app_v2.py:19-45,131-144,
eval/eval_official_widerface.py:160-164, pipeline_v1_detection_first/app_v1.py:44-51 and utils/visualization.py:206-221."""
import math


def choose_slice_params(img_w, img_h, prefer="auto"):
    long_side = max(img_w, img_h)
    if prefer == "3x3" or (prefer == "auto" and long_side < 3000):
        cols, rows = 3, 3
    else:
        cols, rows = 4, 4
    slice_w = math.ceil(img_w / cols)
    slice_h = math.ceil(img_h / rows)

    def round64(x):
        return int(math.ceil(x / 64) * 64)

    return min(round64(slice_h), img_h), min(round64(slice_w), img_w), 0.2, 0.2


def adaptive_slice_size(w, h):
    max_dim = max(w, h)
    if max_dim > 2500:
        return 512
    if max_dim > 1500:
        return 416
    return 320


def detection_first_slice_size(dim, base=512):
    return max(dim // 2 if dim < base * 1.5 else base, 1)


def crop_rectangles(boxes, w, h):
    out = []
    for b in boxes:
        x1, y1, x2, y2 = [int(c) for c in b]
        x1, y1 = max(0, x1), max(0, y1)
        x2, y2 = min(w, x2), min(h, y2)
        if (y2 - y1) * (x2 - x1) > 0 and y2 > y1 and x2 > x1:
            out.append((x1, y1, x2, y2))
    return out
