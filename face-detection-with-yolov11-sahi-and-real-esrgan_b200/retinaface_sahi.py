"""`RetinaFaceSAHI` — mirror of the reference's `docs sahi/retinaface_sahi.py:19-275` (the file BASELINE config 3 names).

The reference class wraps insightface's `FaceAnalysis` (SCRFD / RetinaFace through onnxruntime) as a SAHI plug-in.  Under the
vendored plug-in base (`docs sahi/base.py:162-189`) it has two quirks, both reproduced here by default and pinned by
tests/golden/retinaface_sahi_outputs.json (generated from the reference's own class):

  1. `_create_object_prediction_list_from_original_predictions` RETURNS the list (:187-261) instead of storing it in
     `_object_prediction_list_per_image`, so `convert_original_predictions` leaves `object_prediction_list` empty and
     `get_sliced_prediction` / `get_prediction` see no detections at all;
  2. the returned boxes are already shifted into full-image coordinates (:227-230) AND carry `shift_amount` (:252): calling
     `get_shifted_object_prediction()` on them shifts twice.
  (also: the base constructor resets `image_size` to None, so the detector is always prepared at 640x640.)

`reference_quirks=False` is the usable form: slice-local boxes + `shift_amount`, stored where the base class looks for them —
the behaviour of `utils/insightface_wrapper.py:52-98`, which is the working stand-in for config 3 — and the cross-slice merge
then runs through Kernel 3 inside `get_sliced_prediction`'s generic path.  insightface / onnxruntime are not in this image:
pass `model=<FaceAnalysis-like object>` (`.prepare(ctx_id, det_size, det_thresh)`, `.get(img)` -> faces with `.bbox`,
`.det_score`), or the constructor raises ImportError like the reference's import does."""
from __future__ import annotations

import logging
from typing import List, Optional

import numpy as np

from .sahi_api.base import DetectionModel
from .sahi_api.prediction import ObjectPrediction


def _safe_int(val, default: int) -> int:
    """:9-17 — a positive int / numpy integer or an all-digit string, else the default."""
    if isinstance(val, (int, np.integer)) and not isinstance(val, bool) and val > 0:
        return int(val)
    if isinstance(val, str) and val.isdigit():
        return int(val)
    return int(default)


class RetinaFaceSAHI(DetectionModel):
    def __init__(self, model_path: Optional[str] = None, confidence_threshold: float = 0.5, device: str = "cpu",
                 category_mapping: Optional[dict] = None, load_at_init: bool = True, image_size: Optional[int] = 640,
                 ctx_id: Optional[int] = None, model=None, reference_quirks: bool = True):
        self._raw_predictions = []
        self.reference_quirks = bool(reference_quirks)
        self.ctx_id = -1 if str(device).startswith("cpu") else (0 if ctx_id is None else ctx_id)
        self.image_size = _safe_int(image_size, 640)
        self.logger = logging.getLogger(__name__)
        self._injected = model
        # (the base constructor stores image_size=None over the value above — quirk kept: det_size is always 640)
        super().__init__(model_path=model_path, confidence_threshold=float(confidence_threshold), device=device,
                         category_mapping=category_mapping, load_at_init=load_at_init)

    def set_device(self, device=None):
        self.device = device  # a string, like the reference's sahi build keeps it; the detector itself is onnxruntime's

    def _resolved_det_size(self) -> tuple:
        side = _safe_int(getattr(self, "image_size", None), 640)
        return (side, side)

    def load_model(self):
        if self._injected is not None:
            self.model = self._injected
        else:
            try:
                from insightface.app import FaceAnalysis
            except ImportError as e:
                raise ImportError("insightface is not installed: pass model=<FaceAnalysis-like object>") from e
            providers = ["CPUExecutionProvider"]
            if self.ctx_id >= 0:
                try:
                    import onnxruntime as ort

                    if "CUDAExecutionProvider" in ort.get_available_providers():
                        providers = ["CUDAExecutionProvider", "CPUExecutionProvider"]
                    else:
                        self.ctx_id = -1
                except Exception:
                    self.ctx_id = -1
            self.model = FaceAnalysis(providers=providers)
        self.model.prepare(ctx_id=self.ctx_id, det_size=self._resolved_det_size(),
                           det_thresh=float(getattr(self, "confidence_threshold", 0.5)))

    def unload_model(self):
        if hasattr(self, "model"):
            del self.model

    # ---- inference ------------------------------------------------------------------------------------------------
    def _faces_to_predictions(self, faces, width, height, shift, full_shape, clamp_before_check: bool):
        """Shared conversion: int() truncation of the float box, optional shift, clamp to [0,w]x[0,h], score gate."""
        out: List[ObjectPrediction] = []
        thr = float(getattr(self, "confidence_threshold", 0.5))
        for face in faces:
            bbox = getattr(face, "bbox", None)
            if bbox is None:
                continue
            try:
                x1, y1, x2, y2 = (int(v) for v in bbox)
            except Exception:
                continue
            if clamp_before_check:  # the shifted form clamps first, then drops empty boxes (:227-240)
                x1, y1, x2, y2 = x1 + shift[0], y1 + shift[1], x2 + shift[0], y2 + shift[1]
            elif x2 <= x1 or y2 <= y1:  # the direct form drops empty boxes before clamping (:150-156)
                continue
            x1, x2 = max(0, min(width, x1)), max(0, min(width, x2))
            y1, y2 = max(0, min(height, y1)), max(0, min(height, y2))
            if clamp_before_check and (x2 <= x1 or y2 <= y1):
                continue
            raw = getattr(face, "det_score", None)
            score = float(raw) if raw is not None else 0.0
            if score < thr:
                continue
            kw = dict(shift_amount=shift, full_shape=full_shape) if clamp_before_check else {}
            out.append(ObjectPrediction(bbox=[x1, y1, x2, y2], score=score, category_id=0, category_name="face", **kw))
        return out

    def perform_inference(self, image: np.ndarray):
        """:96-180 — runs the detector, keeps the raw faces for `original_predictions`, and (unlike the plug-in protocol)
        also RETURNS ObjectPredictions in the coordinates of `image`; never raises (returns [])."""
        try:
            if image is None or image.size == 0:
                return []
            if image.dtype != np.uint8:
                image = np.clip(image * 255, 0, 255).astype(np.uint8)
            if image.ndim != 3 or image.shape[2] != 3:
                return []
            ds = getattr(self.model, "det_size", None)
            if ds is None or None in ds:
                self.model.prepare(ctx_id=self.ctx_id, det_size=self._resolved_det_size(),
                                   det_thresh=float(getattr(self, "confidence_threshold", 0.5)))
            faces = self.model.get(image)
            self._raw_predictions = faces
            if len(faces) == 0:
                return []
            h, w = image.shape[:2]
            return self._faces_to_predictions(faces, w, h, (0, 0), None, clamp_before_check=False)
        except Exception as e:  # the reference swallows everything here
            self.logger.warning("perform_inference error: %s", e)
            return []

    @property
    def original_predictions(self):
        return getattr(self, "_raw_predictions", [])

    def _create_object_prediction_list_from_original_predictions(self, shift_amount_list=None, full_shape_list=None):
        """:187-267.  reference_quirks=True: returns pre-shifted predictions and stores nothing (see the module docstring);
        False: stores slice-local predictions with `shift_amount` in `_object_prediction_list_per_image` (and returns them)."""
        faces = self._raw_predictions
        shift = shift_amount_list if isinstance(shift_amount_list, list) and len(shift_amount_list) >= 2 else [0, 0]
        if isinstance(shift[0], (list, tuple)):  # sahi's usual [[x, y]] form
            shift = list(shift[0])
        full = full_shape_list if isinstance(full_shape_list, list) and len(full_shape_list) >= 2 else [1024, 1024]
        if isinstance(full[0], (list, tuple)):
            full = list(full[0])
        if self.reference_quirks:
            if not faces:
                return []
            return self._faces_to_predictions(faces, full[1], full[0], shift, full, clamp_before_check=True)
        out = []
        thr = float(self.confidence_threshold)
        for face in faces or []:
            score = float(face.det_score)
            if score < thr:
                continue
            x1, y1, x2, y2 = (int(v) for v in face.bbox)  # astype(int)-style truncation, utils/insightface_wrapper.py:86
            out.append(ObjectPrediction(bbox=[x1, y1, x2, y2], category_id=0, category_name="face", score=score,
                                        shift_amount=shift, full_shape=full_shape_list if full_shape_list is not None else None))
        self._object_prediction_list_per_image = [out]
        return out

    @property
    def category_names(self) -> List[str]:
        return ["face"]

    @property
    def has_mask(self) -> bool:
        return False

    @property
    def model_name(self) -> str:
        return "RetinaFace"
