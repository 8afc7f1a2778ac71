"""`ultralytics.YOLO`-shaped front end over the B200 pipeline (the `.model` attribute of the reference plugin,
utils/yolo_wrapper.py:55,74-80; also called raw by eval/eval_official_widerface.py:149,219).

`YOLO(path)` loads a file saved by `YOLO.save()` or an ultralytics checkpoint (checkpoints.py reads the pickle without
ultralytics and folds Conv+BN); an unloadable path raises.  `YOLO("random-init")` is the deterministic random-init
YOLO11n-pose of backbones/yolo11_pose.py (there is no network for real checkpoints; parity is judged after the backbone).  `predict()` runs Kernel 1 -> backbone -> Kernel 2a ->
Kernel 3 (per-image NMS) -> Kernel 2b on the GPU and returns `Results` with CUDA tensors."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .backbones.yolo11_pose import YOLO11Pose, build_yolo11n_pose
from .engine import SlicedFaceDetector


class Boxes:
    def __init__(self, xyxy: torch.Tensor, conf: torch.Tensor):
        self.xyxy, self.conf = xyxy, conf
        self.cls = torch.zeros_like(conf)

    @property
    def data(self):
        return torch.cat((self.xyxy, self.conf[:, None], self.cls[:, None]), 1)

    @property
    def xywh(self):
        wh = self.xyxy[:, 2:] - self.xyxy[:, :2]
        return torch.cat((self.xyxy[:, :2] + wh / 2, wh), 1)

    def __len__(self):
        return int(self.xyxy.shape[0])


class Keypoints:
    def __init__(self, data: torch.Tensor):
        self.data = data  # [n,5,3] x, y, conf

    @property
    def xy(self):
        return self.data[..., :2]

    @property
    def conf(self):
        return self.data[..., 2]

    def __len__(self):
        return int(self.data.shape[0])


class Results:
    def __init__(self, boxes: Boxes, keypoints: Keypoints, orig_shape, names):
        self.boxes, self.keypoints, self.orig_shape, self.names = boxes, keypoints, orig_shape, names

    def __len__(self):
        return len(self.boxes)


class YOLO:
    names = {0: "face"}

    def __init__(self, model="yolo11n-pose.pt", task: Optional[str] = None, verbose: bool = False, seed: int = 0,
                 allow_random_init: bool = False):
        """`model`: a checkpoint path (this package's `save()` format or an ultralytics `best.pt`, read without
        ultralytics by checkpoints.read_ultralytics_checkpoint), a torch module, or the literal "random-init".
        Like `ultralytics.YOLO(path)` (utils/yolo_wrapper.py:55) a path that cannot be loaded RAISES — FileNotFoundError
        or checkpoints.CheckpointError; the deterministic random-init YOLO11n-pose (there is no network for real weights
        here: tests, bench) must be requested explicitly with model="random-init" or allow_random_init=True."""
        self.ckpt_path = model if isinstance(model, str) else None
        self.info = dict(scale="n", nc=1, kpt_shape=(5, 3), names=dict(self.names), source="random-init")
        if isinstance(model, torch.nn.Module):
            self.model = model
            self.info["source"] = "module"
        elif model == "random-init":
            self.model = build_yolo11n_pose(seed=seed)
        else:
            from .checkpoints import load_yolo

            try:
                self.model, info = load_yolo(model)
                self.info.update(info, source=model)
                self.names = dict(info.get("names") or self.names)
            except (FileNotFoundError, ValueError):
                if not allow_random_init:
                    raise
                print(f"[fsd_b200] {model}: not loadable, allow_random_init=True -> random-init YOLO11n-pose")
                self.model = build_yolo11n_pose(seed=seed)
        self.task = "pose"
        self._engines = {}

    def save(self, path: str):
        torch.save({"model": self.model.state_dict(), "arch": {k: (list(v) if isinstance(v, tuple) else v)
                                                             for k, v in getattr(self.model, "arch", {}).items()}}, path)

    def to(self, device):
        return self

    def engine(self, device, half: bool = True) -> SlicedFaceDetector:
        key = (str(device), half)
        eng = self._engines.get(key)
        if eng is None:
            eng = SlicedFaceDetector(self.model, device=device, half=half, truncate=False)
            self._engines[key] = eng
        return eng

    @torch.no_grad()
    def predict(self, source=None, conf: float = 0.25, device=None, imgsz: int = 640, verbose: bool = False,
                half: bool = True, **kwargs):
        if isinstance(source, str):
            import cv2

            source = cv2.imread(source)  # ultralytics reads paths with cv2 (BGR)
        if not isinstance(source, np.ndarray) or source.ndim != 3:
            raise TypeError("YOLO.predict expects a path or one HWC uint8 ndarray")
        if device is None or str(device) == "cpu":
            if not torch.cuda.is_available():
                raise RuntimeError("fsd_b200.YOLO runs on CUDA only (no CPU fallback); no CUDA device is visible")
            device = "cuda:0"
        boxes, scores, kpts = self.engine(device, half).predict_array(source, conf=conf, imgsz=imgsz)
        return [Results(Boxes(boxes, scores), Keypoints(kpts), source.shape[:2], self.names)]

    def __call__(self, source=None, **kwargs):
        return self.predict(source=source, **kwargs)
