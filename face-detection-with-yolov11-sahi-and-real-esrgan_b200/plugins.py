"""Detector plug-ins with the reference's class contracts.

`YOLOv11PoseDetectionModel` mirrors utils/yolo_wrapper.py:7-229 (same constructor, attributes and methods, incl. the
`keypoints_cache` side channel and `attach_keypoints_to_predictions`) and adds the batched device path that
`get_sliced_prediction` uses when it recognises this class.  `InsightFaceDetectionModel` mirrors
utils/insightface_wrapper.py:7-113 over any FaceAnalysis-like object (`.get(img)` -> faces with .bbox/.det_score)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .sahi_api.base import DetectionModel
from .sahi_api.prediction import ObjectPrediction
from .yolo import YOLO


def _first(lst, default):
    if lst is None:
        return default
    if isinstance(lst, list) and len(lst) > 0 and isinstance(lst[0], list):
        return lst[0]
    return lst


class YOLOv11PoseDetectionModel(DetectionModel):
    @property
    def supports_batched_slices(self):
        """get_sliced_prediction takes the fused device path when `.model` is the B200 YOLO front end."""
        return isinstance(self.model, YOLO)

    def __init__(self, model_path: str = None, confidence_threshold: float = 0.3, device: str = "cpu",
                 image_size: int = 1024, **kwargs):
        self._model_path, self._device, self._image_size = model_path, device, image_size
        self._confidence_threshold = confidence_threshold
        self.keypoints_cache = {}
        self.half = kwargs.pop("half", True)
        super().__init__(model_path=model_path, confidence_threshold=confidence_threshold, device=device, **kwargs)
        self.model_path, self.device, self.image_size = self._model_path, self._device, self._image_size
        self.confidence_threshold = self._confidence_threshold

    def set_device(self, device=None):
        self.device = device

    def load_model(self):
        if not self.model_path:
            raise ValueError("model_path harus ditentukan")
        self.model = YOLO(self.model_path)
        self.category_mapping = {"0": "face"}

    def set_model(self, model, **kwargs):
        # an fsd_b200.YOLO, a torch backbone to wrap, or any object with ultralytics' `.predict` surface
        if isinstance(model, YOLO) or hasattr(model, "predict"):
            self.model = model
        else:
            self.model = YOLO(model)
        self.category_mapping = {"0": "face"}

    def unload_model(self):
        self.model = None
        self.keypoints_cache = {}

    def _cuda_device(self):
        from . import ops

        return ops.resolve_device(self.device)

    def engine(self):
        eng = self.model.engine(self._cuda_device(), self.half)
        eng.conf, eng.imgsz = self.confidence_threshold, self.image_size
        return eng

    # ---- per-slice protocol (reference semantics, one image at a time) ----------------------------------
    def perform_inference(self, image: np.ndarray):
        if image.dtype != np.uint8:
            image = (image * 255).astype(np.uint8)
        extra = {"half": self.half} if isinstance(self.model, YOLO) else {}
        self._original_predictions = self.model.predict(source=image, conf=self.confidence_threshold,
                                                        device=self._cuda_device(), imgsz=self.image_size,
                                                        verbose=False, **extra)

    def _create_object_prediction_list_from_original_predictions(self, shift_amount_list=[[0, 0]], full_shape_list=None):
        preds = self._original_predictions
        if not preds or len(preds[0].boxes) == 0:
            self._object_prediction_list_per_image = [[]]
            return
        shift = _first(shift_amount_list, [0, 0])
        if not (isinstance(shift, (list, tuple)) and len(shift) == 2):
            shift = [0, 0]
        full_shape = _first(full_shape_list, None)
        res = preds[0]
        xyxy = res.boxes.xyxy.cpu().numpy()  # ONE device->host copy per slice (the reference does two per box)
        conf = res.boxes.conf.cpu().numpy()
        kpts = res.keypoints.data.cpu().numpy() if res.keypoints is not None else None
        out = []
        for i in range(len(xyxy)):
            x1, y1, x2, y2 = (int(v) for v in xyxy[i].astype(int))
            out.append(ObjectPrediction(bbox=[x1, y1, x2, y2], category_id=0, category_name="face",
                                        score=float(conf[i]), shift_amount=shift, full_shape=full_shape))
            if kpts is not None and i < len(kpts):
                k = kpts[i].copy()
                k[:, 0] += shift[0]
                k[:, 1] += shift[1]
                self.keypoints_cache[f"{x1 + shift[0]}_{y1 + shift[1]}_{x2 + shift[0]}_{y2 + shift[1]}"] = k
        self._object_prediction_list_per_image = [out]

    # ---- key-points ---------------------------------------------------------------------------------------
    def attach_keypoints_to_predictions(self, object_prediction_list):
        """Same selection rule as the reference (exact box key, else best IoU > 0.5), evaluated by the
        fsd_attach_keypoints kernel over this plug-in's cache instead of a Python double loop."""
        todo = [p for p in object_prediction_list if not hasattr(p, "keypoints") or p.keypoints is None]
        if not todo or not self.keypoints_cache:
            return object_prediction_list
        from . import ops

        keys = list(self.keypoints_cache.keys())
        dev = torch.device(self._cuda_device())
        dets = torch.tensor([[int(float(v)) for v in k.split("_")] for k in keys], dtype=torch.float32, device=dev)
        merged = torch.tensor([[float(v) for v in p.bbox.to_voc_bbox()] for p in todo], dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        src = ops.attach_keypoints(merged, torch.zeros(1, **i32), torch.tensor([len(todo)], **i32), dets,
                                   torch.zeros(1, **i32), torch.tensor([len(keys)], **i32)).cpu().numpy()
        for p, s in zip(todo, src):
            if s >= 0:
                p.keypoints = self.keypoints_cache[keys[int(s)]]
        return object_prediction_list

    def get_keypoints_for_bbox(self, bbox):
        """pipeline_v4_yolo/app_yolo_sahi.py:80-84 calls this (the reference class lacks it and crashes there)."""
        b = [int(v) for v in bbox]
        return self.keypoints_cache.get(f"{b[0]}_{b[1]}_{b[2]}_{b[3]}")

    @property
    def num_categories(self):
        return len(self.category_names)

    @property
    def has_mask(self):
        return False

    @property
    def category_names(self):
        return ["face"]


class InsightFaceDetectionModel(DetectionModel):
    """utils/insightface_wrapper.py contract; `model` is any object with FaceAnalysis' `.get(image)` (insightface and
    onnxruntime are not in this image, so the detector itself is supplied by the caller / a synthetic stand-in).
    Shift and merge still run on the GPU through get_sliced_prediction's generic path (Kernel 3)."""

    def __init__(self, confidence_threshold: float = 0.3, providers: Optional[List[str]] = None, **kwargs):
        self.providers = providers
        kwargs.pop("device", None)
        super().__init__(confidence_threshold=confidence_threshold, device=None, **kwargs)

    def set_device(self, device=None):
        self.device = device

    def load_model(self):
        try:
            from insightface.app import FaceAnalysis
        except ImportError as e:
            raise ImportError("insightface is not installed: pass model=<FaceAnalysis-like object>") from e
        providers = self.providers or ["CPUExecutionProvider"]
        self.model = FaceAnalysis(providers=providers)
        self.model.prepare(ctx_id=0 if "CUDAExecutionProvider" in providers else -1, det_size=(640, 640),
                           det_thresh=self.confidence_threshold)
        self.category_mapping = {"0": "face"}

    def set_model(self, model, **kwargs):
        self.model = model
        self.category_mapping = {"0": "face"}

    def unload_model(self):
        self.model = None

    def perform_inference(self, image: np.ndarray):
        if image.dtype != np.uint8:
            image = (image * 255).astype(np.uint8)
        self._original_predictions = self.model.get(image)

    def _create_object_prediction_list_from_original_predictions(self, shift_amount_list=[[0, 0]], full_shape_list=None):
        faces = self._original_predictions
        if not faces:
            self._object_prediction_list_per_image = [[]]
            return
        out = []
        for face in faces:
            score = float(face.det_score)
            if score < self.confidence_threshold:
                continue
            x1, y1, x2, y2 = face.bbox.astype(int)
            out.append(ObjectPrediction(bbox=[x1, y1, x2, y2], category_id=0, category_name="face", score=score,
                                        shift_amount=shift_amount_list, full_shape=full_shape_list))
        self._object_prediction_list_per_image = [out]

    @property
    def num_categories(self):
        return len(self.category_names)

    @property
    def has_mask(self):
        return False

    @property
    def category_names(self):
        return ["face"]
