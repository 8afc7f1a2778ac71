"""sahi.models.base mirror (reference: docs sahi/base.py:12-196) — the detector plug-in protocol."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from .annotation import Category


class DetectionModel:
    required_packages: List[str] = []

    def __init__(self, model_path: Optional[str] = None, model: Optional[Any] = None, config_path: Optional[str] = None,
                 device: Optional[str] = None, mask_threshold: float = 0.5, confidence_threshold: float = 0.3,
                 category_mapping: Optional[Dict] = None, category_remapping: Optional[Dict] = None,
                 load_at_init: bool = True, image_size: Optional[int] = None):
        self.model_path = model_path
        self.config_path = config_path
        self.model = None
        self.mask_threshold = mask_threshold
        self.confidence_threshold = confidence_threshold
        self.category_mapping = category_mapping
        self.category_remapping = category_remapping
        self.image_size = image_size
        self._original_predictions = None
        self._object_prediction_list_per_image = None
        self.set_device(device)
        self.check_dependencies()
        if load_at_init:
            if model:
                self.set_model(model)
            else:
                self.load_model()

    def check_dependencies(self, packages: Optional[List[str]] = None) -> None:
        import importlib

        for pkg in (packages if packages is not None else getattr(self, "required_packages", [])):
            importlib.import_module(pkg)

    def load_model(self):
        raise NotImplementedError()

    def set_model(self, model: Any, **kwargs):
        raise NotImplementedError()

    def set_device(self, device: Optional[str] = None):
        import torch

        if device is None:
            device = "cuda:0" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device) if not isinstance(device, torch.device) else device

    def unload_model(self):
        import torch

        self.model = None
        if torch.cuda.is_available():
            torch.cuda.empty_cache()

    def perform_inference(self, image: np.ndarray):
        raise NotImplementedError()

    def _create_object_prediction_list_from_original_predictions(self, shift_amount_list=[[0, 0]], full_shape_list=None):
        raise NotImplementedError()

    def _apply_category_remapping(self):
        if self.category_remapping is None:
            raise ValueError("self.category_remapping cannot be None")
        if not isinstance(self._object_prediction_list_per_image, list):
            return
        for per_image in self._object_prediction_list_per_image:
            for op in per_image:
                op.category = Category(id=self.category_remapping[str(op.category.id)], name=op.category.name)

    def convert_original_predictions(self, shift_amount=[[0, 0]], full_shape=None):
        self._create_object_prediction_list_from_original_predictions(shift_amount_list=shift_amount,
                                                                      full_shape_list=full_shape)
        if self.category_remapping:
            self._apply_category_remapping()

    @property
    def object_prediction_list(self):
        per_image = self._object_prediction_list_per_image
        return per_image[0] if per_image else []

    @property
    def object_prediction_list_per_image(self):
        return self._object_prediction_list_per_image or []

    @property
    def original_predictions(self):
        return self._original_predictions
