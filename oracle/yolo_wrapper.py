"""Oracle plug-ins (test infrastructure, see oracle/__init__): CPU restatements of the reference's detector wrappers
utils/yolo_wrapper.py:7-229 (YOLOv11PoseDetectionModel) and utils/insightface_wrapper.py:7-113, over oracle.predict's
DetectionModel.  Pinned: tests/golden holds outputs of the reference's own classes imported unmodified."""
from __future__ import annotations

import numpy as np

from .annotation import ObjectPrediction
from .predict import DetectionModel


class YOLOv11PoseDetectionModel(DetectionModel):
    def __init__(self, model_path=None, confidence_threshold=0.3, device="cpu", image_size=1024, **kwargs):
        self.keypoints_cache = {}
        super().__init__(model_path=model_path, confidence_threshold=confidence_threshold, device=device, **kwargs)
        self.model_path, self.device, self.image_size = model_path, device, image_size
        self.confidence_threshold = confidence_threshold

    def load_model(self):
        raise ValueError("the oracle plug-in is built with model=<OracleYOLO>")

    def set_model(self, model, **kwargs):
        self.model = model
        self.category_mapping = {"0": "face"}

    def unload_model(self):
        self.model = None
        self.keypoints_cache = {}

    def perform_inference(self, image):  # :63-82
        if image.dtype != np.uint8:
            image = (image * 255).astype(np.uint8)
        self._original_predictions = self.model.predict(source=image, conf=self.confidence_threshold,
                                                        device=self.device, imgsz=self.image_size, verbose=False)

    def _create_object_prediction_list_from_original_predictions(self, shift_amount_list=[[0, 0]], full_shape_list=None):
        preds = self._original_predictions  # :84-166
        if not preds or len(preds[0].boxes) == 0:
            self._object_prediction_list_per_image = [[]]
            return
        if shift_amount_list is None:
            shift = [0, 0]
        elif isinstance(shift_amount_list, list):
            if len(shift_amount_list) > 0 and isinstance(shift_amount_list[0], list):
                shift = shift_amount_list[0]
            else:
                shift = shift_amount_list if len(shift_amount_list) == 2 else [0, 0]
        else:
            shift = [0, 0]
        if full_shape_list is None:
            full_shape = None
        elif isinstance(full_shape_list, list):
            full_shape = full_shape_list[0] if len(full_shape_list) > 0 and isinstance(full_shape_list[0], list) else full_shape_list
        else:
            full_shape = None
        result = preds[0]
        boxes = result.boxes
        kdata = result.keypoints.data.cpu().numpy() if getattr(result, "keypoints", None) is not None else None
        out = []
        for i in range(len(boxes)):
            score = float(boxes.conf[i])
            x1, y1, x2, y2 = boxes.xyxy[i].cpu().numpy().astype(int)  # truncation toward zero (:138)
            out.append(ObjectPrediction(bbox=[int(x1), int(y1), int(x2), int(y2)], category_id=0, category_name="face",
                                        score=score, shift_amount=shift, full_shape=full_shape))
            if kdata is not None and i < len(kdata):
                k = kdata[i].copy()
                k[:, 0] += shift[0]
                k[:, 1] += shift[1]
                self.keypoints_cache[f"{int(x1) + shift[0]}_{int(y1) + shift[1]}_{int(x2) + shift[0]}_{int(y2) + shift[1]}"] = k
        self._object_prediction_list_per_image = [out]

    def attach_keypoints_to_predictions(self, object_prediction_list):  # :168-200
        for pred in object_prediction_list:
            bbox = pred.bbox.to_voc_bbox()
            key = f"{bbox[0]}_{bbox[1]}_{bbox[2]}_{bbox[3]}"
            if key in self.keypoints_cache:
                pred.keypoints = self.keypoints_cache[key]
                continue
            best_iou, best = 0.0, None
            for k, kpts in self.keypoints_cache.items():
                iou = self._calculate_iou(bbox, [int(float(x)) for x in k.split("_")])
                if iou > best_iou:
                    best_iou, best = iou, kpts
            if best_iou > 0.5 and best is not None:
                pred.keypoints = best
        return object_prediction_list

    @staticmethod
    def _calculate_iou(b1, b2):  # :202-217
        x1, y1, x2, y2 = max(b1[0], b2[0]), max(b1[1], b2[1]), min(b1[2], b2[2]), min(b1[3], b2[3])
        if x2 < x1 or y2 < y1:
            return 0.0
        inter = (x2 - x1) * (y2 - y1)
        union = (b1[2] - b1[0]) * (b1[3] - b1[1]) + (b2[2] - b2[0]) * (b2[3] - b2[1]) - inter
        return inter / union if union > 0 else 0.0

    num_categories = property(lambda self: 1)
    has_mask = property(lambda self: False)
    category_names = property(lambda self: ["face"])


class InsightFaceDetectionModel(DetectionModel):  # utils/insightface_wrapper.py:7-113
    def __init__(self, confidence_threshold=0.3, providers=None, **kwargs):
        self.providers = providers
        kwargs.pop("device", None)
        super().__init__(confidence_threshold=confidence_threshold, device=None, **kwargs)

    def load_model(self):
        raise ValueError("the oracle plug-in is built with model=<FaceAnalysis-like object>")

    def set_model(self, model, **kwargs):
        self.model = model
        self.category_mapping = {"0": "face"}

    def perform_inference(self, image):
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

    num_categories = property(lambda self: 1)
    has_mask = property(lambda self: False)
    category_names = property(lambda self: ["face"])
