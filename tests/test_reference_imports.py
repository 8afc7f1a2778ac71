"""Boundary test (CPU, only where the reference checkout exists): the reference's own files import UNMODIFIED on top of
compat/ and expose the classes its apps use; their CUDA-free surface behaves."""
import importlib.util
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("FSD_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this machine")


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def compat():
    import fsd_b200

    d = fsd_b200.install_compat()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("sahi", "ultralytics", "realesrgan", "basicsr", "bbox")}
    sys.path.insert(0, REF)
    yield d
    sys.path.remove(REF)
    for k in list(sys.modules):
        if k.split(".")[0] in ("sahi", "ultralytics", "realesrgan", "basicsr", "bbox", "utils"):
            sys.modules.pop(k)
    sys.modules.update(saved)


def test_reference_plugin_and_enhancer_import_unmodified(compat, tmp_path):
    yw = _load("ref_yolo_wrapper_compat", "utils/yolo_wrapper.py")
    from fsd_b200.sahi_api import DetectionModel
    from fsd_b200.yolo import YOLO

    assert issubclass(yw.YOLOv11PoseDetectionModel, DetectionModel) and yw.YOLO is YOLO
    # like ultralytics.YOLO(path), a missing checkpoint raises out of the reference's load_model (utils/yolo_wrapper.py:55)
    with pytest.raises(FileNotFoundError):
        yw.YOLOv11PoseDetectionModel(model_path="weights-that-do-not-exist.pt", confidence_threshold=0.6, device="cuda:0")
    ckpt = str(tmp_path / "best.pt")
    YOLO("random-init").save(ckpt)
    m = yw.YOLOv11PoseDetectionModel(model_path=ckpt, confidence_threshold=0.6, device="cuda:0")
    assert isinstance(m.model, YOLO) and m.image_size == 1024 and m.category_mapping == {"0": "face"}
    assert m.category_names == ["face"] and m.has_mask is False and m.keypoints_cache == {}
    from fsd_b200.sahi_api.predict import _fused_capable

    assert _fused_capable(m)  # the unmodified reference class takes the fused device path
    enh = _load("ref_enhancer_compat", "utils/enhancer.py")
    from fsd_b200.backbones.rrdbnet import RRDBNet
    from fsd_b200.enhancer import RealESRGANer

    assert enh.RRDBNet is RRDBNet and enh.RealESRGANer is RealESRGANer and hasattr(enh.FaceEnhancer, "enhance_image")
    iw = pytest.importorskip  # insightface itself is absent: the wrapper's import must fail on THAT, not on sahi
    with pytest.raises(ModuleNotFoundError, match="insightface"):
        _load("ref_insight_compat", "utils/insightface_wrapper.py")


def test_reference_cli_and_evaluator_import_unmodified(compat, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    app = _load("ref_app_yolo_inference", "pipeline_v4_yolo/app_yolo_inference.py")
    from fsd_b200.sahi_api.predict import get_sliced_prediction

    assert app.get_sliced_prediction is get_sliced_prediction and callable(app.main)
    vis = sys.modules["utils.visualization"]
    from fsd_b200.sahi_api import ObjectPrediction

    op = ObjectPrediction(bbox=[4, 5, 30, 40], category_id=0, category_name="face", score=0.9)
    op.keypoints = np.array([[10, 10, 0.9]] * 5, dtype=np.float32)
    import cv2

    from fsd_b200.sahi_api import PredictionResult

    cv2.imwrite("in.png", np.zeros((64, 64, 3), np.uint8))
    res = PredictionResult([op], "in.png", {})
    vis.draw_detections("in.png", res, "out/vis.png")  # consumes bbox / score / keypoints of OUR objects
    out = cv2.imread("out/vis.png")
    assert out is not None and out.any()
    vis.save_face_crops("in.png", res, "out/crops") if hasattr(vis, "save_face_crops") else None
