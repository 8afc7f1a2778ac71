"""Oracle plug-in (test infrastructure, see oracle/__init__): CPU restatement of `docs sahi/retinaface_sahi.py:19-275`
(`RetinaFaceSAHI`) over oracle.predict's DetectionModel, with the reference's quirks as they are: the conversion method
returns pre-shifted predictions and stores nothing, so the SAHI driver sees no detections.  Pinned by
tests/golden/retinaface_sahi_outputs.json (the reference's own class, imported unmodified)."""
from __future__ import annotations

import numpy as np

from .annotation import ObjectPrediction
from .predict import DetectionModel


def safe_int(val, default):
    if isinstance(val, (int, np.integer)) and not isinstance(val, bool) and val > 0:
        return int(val)
    if isinstance(val, str) and val.isdigit():
        return int(val)
    return int(default)


class RetinaFaceSAHI(DetectionModel):
    def __init__(self, model=None, model_path=None, confidence_threshold=0.5, device="cpu", category_mapping=None,
                 load_at_init=True, image_size=640, ctx_id=None):
        self._raw_predictions = []
        self.ctx_id = -1 if str(device).startswith("cpu") else (0 if ctx_id is None else ctx_id)
        self.image_size = safe_int(image_size, 640)
        super().__init__(model_path=model_path, model=model, confidence_threshold=float(confidence_threshold), device=device,
                         category_mapping=category_mapping, load_at_init=load_at_init)  # resets image_size to None

    def load_model(self):
        raise ValueError("the oracle plug-in is built with model=<FaceAnalysis-like object>")

    def set_model(self, model, **kwargs):
        self.model = model
        self.model.prepare(ctx_id=self.ctx_id, det_size=self.det_size(), det_thresh=float(self.confidence_threshold))

    def det_size(self):
        side = safe_int(getattr(self, "image_size", None), 640)
        return (side, side)

    def perform_inference(self, image):  # :96-180
        if image is None or image.size == 0:
            return []
        if image.dtype != np.uint8:
            image = np.clip(image * 255, 0, 255).astype(np.uint8)
        if image.ndim != 3 or image.shape[2] != 3:
            return []
        faces = self.model.get(image)
        self._raw_predictions = faces
        h, w = image.shape[:2]
        out = []
        for face in faces:
            x1, y1, x2, y2 = [int(v) for v in face.bbox]
            if x2 <= x1 or y2 <= y1:
                continue
            x1 = max(0, min(w, x1)); y1 = max(0, min(h, y1)); x2 = max(0, min(w, x2)); y2 = max(0, min(h, y2))
            score = float(face.det_score)
            if score < float(self.confidence_threshold):
                continue
            out.append(ObjectPrediction(bbox=[x1, y1, x2, y2], score=score, category_id=0, category_name="face"))
        return out

    @property
    def original_predictions(self):
        return getattr(self, "_raw_predictions", [])

    def _create_object_prediction_list_from_original_predictions(self, shift_amount_list=None, full_shape_list=None):  # :187-267
        faces = self._raw_predictions
        if not faces:
            return []
        shift = shift_amount_list if isinstance(shift_amount_list, list) and len(shift_amount_list) >= 2 else [0, 0]
        full = full_shape_list if isinstance(full_shape_list, list) and len(full_shape_list) >= 2 else [1024, 1024]
        h, w = full[0], full[1]
        out = []
        for face in faces:
            x1, y1, x2, y2 = [int(v) for v in face.bbox]
            x1 += shift[0]; x2 += shift[0]; y1 += shift[1]; y2 += shift[1]
            x1 = max(0, min(w, x1)); y1 = max(0, min(h, y1)); x2 = max(0, min(w, x2)); y2 = max(0, min(h, y2))
            if x2 <= x1 or y2 <= y1:
                continue
            score = float(face.det_score)
            if score < self.confidence_threshold:
                continue
            out.append(ObjectPrediction(bbox=[x1, y1, x2, y2], score=score, category_id=0, category_name="face",
                                        shift_amount=shift, full_shape=full))
        return out  # NOT stored: docs sahi/base.py:178-181 ignores the return value

    category_names = property(lambda self: ["face"])
    has_mask = property(lambda self: False)
    model_name = property(lambda self: "RetinaFace")
