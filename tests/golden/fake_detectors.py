"""Deterministic, machine-independent stand-ins for the detector libraries (ultralytics.YOLO, insightface FaceAnalysis)
used by the golden fixtures and the boundary tests.  The test image encodes its own pixel coordinates, so a detector
handed a slice can recover the slice origin and emit boxes for the known synthetic faces it overlaps — with float
coordinates, jitter and scores that depend on (face, slice), like real per-slice detections do."""
from __future__ import annotations

import numpy as np
import torch


def coordinate_image(height: int, width: int) -> np.ndarray:
    """HWC uint8 image whose pixel (y,x) stores (x % 256, y % 256, 16*(x // 256) + (y // 256))."""
    yy, xx = np.mgrid[0:height, 0:width]
    return np.stack([xx % 256, yy % 256, 16 * (xx // 256) + (yy // 256)], -1).astype(np.uint8)


def decode_origin(arr: np.ndarray):
    p = arr[0, 0].astype(int)
    return int(p[0] + 256 * (p[2] // 16)), int(p[1] + 256 * (p[2] % 16))


def synthetic_faces(height: int, width: int, n: int, seed: int):
    rng = np.random.default_rng(seed)
    faces = []
    for _ in range(n):
        w = float(np.exp(rng.uniform(np.log(10), np.log(min(width, height) / 3))))
        h = w * rng.uniform(1.0, 1.4)
        x, y = rng.uniform(0, width - w), rng.uniform(0, height - h)
        faces.append((x, y, x + w, y + h))
    return faces


def _hash01(*ints) -> float:
    v = 1469598103934665603
    for i in ints:
        v = ((v ^ (int(i) & 0xFFFFFFFF)) * 1099511628211) % (1 << 64)
    return (v >> 11) / float(1 << 53)


def detections_for_window(faces, ox, oy, w, h, min_visible=0.3):
    """Float boxes (window-local), scores and 5 key-points for every face visible in the window [ox,oy,ox+w,oy+h)."""
    boxes, scores, kpts = [], [], []
    for fid, (x1, y1, x2, y2) in enumerate(faces):
        ix1, iy1, ix2, iy2 = max(x1, ox), max(y1, oy), min(x2, ox + w), min(y2, oy + h)
        if ix2 <= ix1 or iy2 <= iy1:
            continue
        vis = (ix2 - ix1) * (iy2 - iy1) / ((x2 - x1) * (y2 - y1))
        if vis < min_visible:
            continue
        j = [(_hash01(fid, ox, oy, k) - 0.5) * 4.0 for k in range(4)]
        b = [min(max(ix1 - ox + j[0], 0.0), w), min(max(iy1 - oy + j[1], 0.0), h),
             min(max(ix2 - ox + j[2], 0.0), w), min(max(iy2 - oy + j[3], 0.0), h)]
        if b[2] - b[0] < 1 or b[3] - b[1] < 1:
            continue
        boxes.append(b)
        scores.append(0.35 + 0.64 * _hash01(fid, ox, oy, 99) * vis)
        kp = [[b[0] + (b[2] - b[0]) * fx, b[1] + (b[3] - b[1]) * fy, 0.5 + 0.5 * _hash01(fid, k)]
              for k, (fx, fy) in enumerate([(0.3, 0.35), (0.7, 0.35), (0.5, 0.55), (0.35, 0.75), (0.65, 0.75)])]
        kpts.append(kp)
    order = np.argsort(-np.array(scores), kind="stable") if scores else []
    return ([boxes[i] for i in order], [scores[i] for i in order], [kpts[i] for i in order])


class _T:
    """tiny stand-in for ultralytics Boxes / Keypoints holding CPU tensors"""


class FakeResults:
    def __init__(self, boxes, scores, kpts):
        self.boxes = _T()
        self.boxes.xyxy = torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4)
        self.boxes.conf = torch.tensor(scores, dtype=torch.float32)
        self.boxes.__class__ = type("Boxes", (), {"__len__": lambda s: int(s.xyxy.shape[0])})
        self.keypoints = _T()
        self.keypoints.data = torch.tensor(kpts, dtype=torch.float32).reshape(-1, 5, 3)


class FakeYOLO:
    """`ultralytics.YOLO` surface: predict(source=ndarray, conf, device, imgsz, verbose) -> [Results]."""

    faces = []

    def __init__(self, model_path=None):
        self.model_path = model_path

    def predict(self, source=None, conf=0.25, device=None, imgsz=640, verbose=False, **_):
        ox, oy = decode_origin(source)
        h, w = source.shape[:2]
        boxes, scores, kpts = detections_for_window(type(self).faces, ox, oy, w, h)
        keep = [i for i, s in enumerate(scores) if s > conf]
        return [FakeResults([boxes[i] for i in keep], [scores[i] for i in keep], [kpts[i] for i in keep])]

    __call__ = predict


class FakeFace:
    def __init__(self, bbox, score):
        self.bbox = np.array(bbox, dtype=np.float32)
        self.det_score = np.float32(score)


class FakeFaceAnalysis:
    """`insightface.app.FaceAnalysis` surface: prepare(...), get(img) -> faces with .bbox (float32) and .det_score."""

    faces = []

    def __init__(self, providers=None, **_):
        self.providers = providers

    def prepare(self, ctx_id=0, det_size=(640, 640), det_thresh=0.5):
        self.det_thresh = det_thresh

    def get(self, img):
        ox, oy = decode_origin(img)
        h, w = img.shape[:2]
        boxes, scores, _ = detections_for_window(type(self).faces, ox, oy, w, h)
        return [FakeFace(b, s) for b, s in zip(boxes, scores)]


class AffineUpsampler(torch.nn.Module):
    """Stand-in for basicsr RRDBNet with the same constructor: nearest up-sampling by `scale` and an affine map that
    leaves [0,1] on both sides.  Exact in fp32 on every machine, so enhancer goldens are bit-reproducible."""

    def __init__(self, num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale = scale
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return torch.nn.functional.interpolate(x, scale_factor=self.scale, mode="nearest") * 1.25 - 0.125
