"""Oracle for Kernel 2 / stage-1 of Kernel 3: ultralytics pose-head decode, NMS and rescale (test infrastructure).

Restates [EXT ultralytics] `Detect/Pose._inference`, `DFL`, `make_anchors`, `dist2bbox`, `kpts_decode`,
`ops.non_max_suppression`, `ops.scale_boxes`, `ops.scale_coords`, `clip_boxes`, and the `YOLO.predict` /
`Results` surface the reference uses (utils/yolo_wrapper.py:74-80,119-137; eval/eval_official_widerface.py:149,219),
following SURVEY.md App. A.4.  Straight-line torch-CPU fp32; per-slice NMS is the real torchvision.ops.nms.
Parity unpinned for the restated parts (ultralytics is absent and unpinned in requirements.txt).
"""
from __future__ import annotations

import numpy as np
import torch
import torchvision

from . import letterbox as olb

STRIDES = (8, 16, 32)
REG_MAX = 16


def make_anchors(level_hw, strides=STRIDES, offset=0.5):
    pts, st = [], []
    for (h, w), s in zip(level_hw, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32))
    return torch.cat(pts).transpose(0, 1), torch.cat(st).transpose(0, 1)  # [2,A], [1,A]


def decode_head(levels, nc=1, kpt_shape=(5, 3)):
    """levels: [(box [B,64,h,w], cls [B,nc,h,w], kpt [B,15,h,w])] -> y [B, 4+nc+15, A] (xywh, sigmoid cls, kpts)."""
    B = levels[0][0].shape[0]
    level_hw = [tuple(l[0].shape[2:]) for l in levels]
    box = torch.cat([l[0].float().reshape(B, 4 * REG_MAX, -1) for l in levels], 2)
    cls = torch.cat([l[1].float().reshape(B, nc, -1) for l in levels], 2)
    kpt = torch.cat([l[2].float().reshape(B, kpt_shape[0] * kpt_shape[1], -1) for l in levels], 2)
    anchors, strides = make_anchors(level_hw)
    A = box.shape[2]
    # DFL: softmax over the 16 bins, expectation with arange(16) (a 1x1 conv with fixed weights upstream)
    prob = box.view(B, 4, REG_MAX, A).transpose(2, 1).softmax(1)  # [B,16,4,A]
    dist = (prob * torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)).sum(1)  # [B,4,A]
    lt, rb = dist.chunk(2, 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides
    # key-points
    ndim = kpt_shape[1]
    y = kpt.view(B, kpt_shape[0], ndim, A).clone()
    a = y[:, :, :2] * 2.0 + (anchors.view(1, 1, 2, A) - 0.5)
    a = a * strides.view(1, 1, 1, A)
    if ndim == 3:
        a = torch.cat((a, y[:, :, 2:3].sigmoid()), 2)
    pk = a.reshape(B, kpt_shape[0] * ndim, A)
    return torch.cat((dbox, cls.sigmoid(), pk), 1)


def xywh2xyxy(x):
    y = torch.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def non_max_suppression(pred, conf_thres=0.25, iou_thres=0.7, max_det=300, nc=1, max_nms=30000, max_wh=7680,
                        return_candidates=False):
    """pred [B, 4+nc+nm, A] -> list of [n, 6+nm] rows (xyxy, conf, cls, extras), score-descending; time limit disabled."""
    B = pred.shape[0]
    mi = 4 + nc
    xc = pred[:, 4:mi].amax(1) > conf_thres
    pred = pred.transpose(-1, -2).clone()
    pred[..., :4] = xywh2xyxy(pred[..., :4])
    out, cands = [], []
    for xi in range(B):
        x = pred[xi][xc[xi]]
        if not x.shape[0]:
            out.append(torch.zeros((0, 6 + pred.shape[2] - mi)))
            cands.append((x, torch.zeros(0, dtype=torch.long)))
            continue
        box, cls, mask = x.split((4, nc, x.shape[1] - mi), 1)
        conf, j = cls.max(1, keepdim=True)
        x = torch.cat((box, conf, j.float(), mask), 1)[conf.view(-1) > conf_thres]
        if x.shape[0] > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * max_wh
        keep = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        cands.append((x, keep))
        out.append(x[keep])
    return (out, cands) if return_candidates else out


def clip_boxes(boxes, shape):
    boxes[..., 0] = boxes[..., 0].clamp(0, shape[1])
    boxes[..., 1] = boxes[..., 1].clamp(0, shape[0])
    boxes[..., 2] = boxes[..., 2].clamp(0, shape[1])
    boxes[..., 3] = boxes[..., 3].clamp(0, shape[0])
    return boxes


def scale_boxes(img1_shape, boxes, img0_shape):
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    boxes[..., 0] -= pad_x
    boxes[..., 1] -= pad_y
    boxes[..., 2] -= pad_x
    boxes[..., 3] -= pad_y
    boxes[..., :4] /= gain
    return clip_boxes(boxes, img0_shape)


def scale_coords(img1_shape, coords, img0_shape):
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = (img1_shape[1] - img0_shape[1] * gain) / 2
    pad_y = (img1_shape[0] - img0_shape[0] * gain) / 2
    coords[..., 0] -= pad_x
    coords[..., 1] -= pad_y
    coords[..., 0] /= gain
    coords[..., 1] /= gain
    coords[..., 0] = coords[..., 0].clamp(0, img0_shape[1])
    coords[..., 1] = coords[..., 1].clamp(0, img0_shape[0])
    return coords


class _Boxes:
    def __init__(self, data):
        self.data = data  # [n,6] xyxy, conf, cls

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, 4]

    @property
    def cls(self):
        return self.data[:, 5]

    def __len__(self):
        return self.data.shape[0]


class _Keypoints:
    def __init__(self, data):
        self.data = data  # [n,5,3]

    def __len__(self):
        return self.data.shape[0]


class Results:
    def __init__(self, boxes, keypoints, orig_shape):
        self.boxes = _Boxes(boxes)
        self.keypoints = _Keypoints(keypoints)
        self.orig_shape = orig_shape


class OracleYOLO:
    """CPU stand-in for ultralytics.YOLO over a PyTorch backbone that returns raw per-level head tensors.

    `head_hook(input_tensor, levels) -> levels` lets parity tests substitute the head tensors the GPU backbone
    produced for the same slice, so that everything after the backbone is compared on identical inputs.
    """

    def __init__(self, backbone, nc=1, kpt_shape=(5, 3), half=False, head_hook=None, stride=32):
        self.backbone, self.nc, self.kpt_shape, self.half, self.head_hook, self.stride = backbone, nc, kpt_shape, half, head_hook, stride
        self.last = {}

    @torch.no_grad()
    def predict(self, source=None, conf=0.25, device=None, imgsz=640, verbose=False, iou=0.7, max_det=300, **_):
        img = source
        assert isinstance(img, np.ndarray) and img.ndim == 3, "oracle YOLO takes one HWC uint8 ndarray"
        im = olb.preprocess(img, imgsz=imgsz, stride=self.stride, half=self.half)
        levels = self.backbone(im.float()) if self.head_hook is None else None
        if self.head_hook is not None:
            levels = self.head_hook(im, levels)
        y = decode_head(levels, self.nc, self.kpt_shape)
        dets, cands = non_max_suppression(y, conf, iou, max_det, nc=self.nc, return_candidates=True)
        d = dets[0]
        net_shape = im.shape[2:]
        boxes = d[:, :6].clone()
        boxes[:, :4] = scale_boxes(net_shape, boxes[:, :4], img.shape[:2])
        kpts = d[:, 6:].reshape(-1, *self.kpt_shape).clone()
        kpts = scale_coords(net_shape, kpts, img.shape[:2])
        self.last = dict(input=im, levels=levels, decoded=y, candidates=cands[0][0], keep=cands[0][1], nms_out=d)
        return [Results(boxes, kpts, img.shape[:2])]

    def __call__(self, source=None, **kw):
        kw.setdefault("imgsz", 640)
        return self.predict(source=source, **kw)
