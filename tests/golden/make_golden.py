"""Generate tests/golden/*.json by running the REFERENCE'S OWN files, imported unmodified from /root/reference:

    docs sahi/predict.py, docs sahi/prediction.py, docs sahi/base.py      (vendored SAHI driver + data classes)
    utils/yolo_wrapper.py, utils/insightface_wrapper.py, utils/enhancer.py (the reference's plug-ins / enhancer)

Their un-vendored imports (sahi.slicing, sahi.postprocess.combine, sahi.annotation, ultralytics, insightface,
realesrgan, basicsr) are satisfied by thin in-memory shims bound to the CPU oracle and to the deterministic fake
detectors of fake_detectors.py.  Run here (the container with /root/reference):  python tests/golden/make_golden.py
"""
from __future__ import annotations

import importlib.util
import json
import logging
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = os.environ.get("FSD_REFERENCE", "/root/reference")

import numpy as np  # noqa: E402

import fake_detectors as fd  # noqa: E402
from oracle import annotation as oann  # noqa: E402
from oracle import esrgan as oesr  # noqa: E402
from oracle import postprocess as opp  # noqa: E402
from oracle import slicing as osl  # noqa: E402


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    """sys.modules entries for everything the reference files import but do not vendor."""
    pkg = _mod("sahi")
    pkg.__path__ = []
    _mod("sahi.logger", logger=logging.getLogger("sahi"))
    _mod("sahi.utils").__path__ = []
    _mod("sahi.utils.import_utils", is_available=lambda name: importlib.util.find_spec(name) is not None,
         check_requirements=lambda pkgs: None)
    _mod("sahi.utils.torch_utils", empty_cuda_cache=lambda: None, select_device=lambda d=None: d if d is not None else "cpu")
    _mod("sahi.utils.cv", IMAGE_EXTENSIONS=[".jpg", ".jpeg", ".png"], VIDEO_EXTENSIONS=[".mp4"], cv2=__import__("cv2"),
         crop_object_predictions=None, get_video_reader=None, read_image_as_pil=osl.read_image_as_pil,
         visualize_object_predictions=None)
    from pathlib import Path

    _mod("sahi.utils.file", Path=Path, increment_path=None, list_files=None, save_json=None, save_pickle=None)
    _mod("sahi.utils.coco", Coco=None, CocoImage=None, CocoPrediction=None)
    _mod("sahi.annotation", BoundingBox=oann.BoundingBox, Category=oann.Category, ObjectAnnotation=oann.ObjectAnnotation)
    _mod("sahi.slicing", slice_image=osl.slice_image, get_slice_bboxes=osl.get_slice_bboxes)
    _mod("sahi.postprocess").__path__ = []
    _mod("sahi.postprocess.combine", GreedyNMMPostprocess=opp.GreedyNMMPostprocess, LSNMSPostprocess=opp.LSNMSPostprocess,
         NMMPostprocess=opp.NMMPostprocess, NMSPostprocess=opp.NMSPostprocess, PostprocessPredictions=opp.PostprocessPredictions)
    _mod("sahi.auto_model", AutoDetectionModel=None)
    _mod("sahi.models").__path__ = []
    _mod("ultralytics", YOLO=fd.FakeYOLO)
    _mod("insightface").__path__ = []
    _mod("insightface.app", FaceAnalysis=fd.FakeFaceAnalysis)
    _mod("basicsr").__path__ = []
    _mod("basicsr.archs").__path__ = []
    _mod("basicsr.archs.rrdbnet_arch", RRDBNet=fd.AffineUpsampler)
    _mod("realesrgan", RealESRGANer=oesr.RealESRGANer)


def load_reference(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_all():
    install_shims()
    pred = load_reference("sahi.prediction", "docs sahi/prediction.py")
    # ObjectPrediction must be the class the oracle merge code instantiates too
    opp.ObjectPrediction = pred.ObjectPrediction
    base = load_reference("sahi.models.base", "docs sahi/base.py")
    _mod("sahi.models.ultralytics", UltralyticsDetectionModel=type("UltralyticsDetectionModel", (), {}))
    predict = load_reference("sahi.predict", "docs sahi/predict.py")
    yolo_wrapper = load_reference("ref_yolo_wrapper", "utils/yolo_wrapper.py")
    insight = load_reference("ref_insightface_wrapper", "utils/insightface_wrapper.py")
    enhancer = load_reference("ref_enhancer", "utils/enhancer.py")
    return predict, pred, base, yolo_wrapper, insight, enhancer


SCENES = [  # (name, H, W, n_faces, seed, slice, overlap, postprocess, metric, thr, conf)
    ("c1_like_greedynmm_ios", 540, 960, 40, 1, 320, 0.2, "GREEDYNMM", "IOS", 0.5, 0.5),
    ("c2_like_nms_ios", 384, 512, 25, 2, 256, 0.2, "NMS", "IOS", 0.5, 0.4),
    ("streamlit_like_overlap025", 600, 800, 30, 3, 320, 0.25, "GREEDYNMM", "IOS", 0.5, 0.5),
    ("nmm_iou", 480, 640, 30, 4, 256, 0.2, "NMM", "IOU", 0.5, 0.45),
    ("small_image_single_slice", 200, 300, 6, 5, 320, 0.2, "GREEDYNMM", "IOS", 0.5, 0.4),
    ("tuning_greedynmm_iou_03", 480, 640, 30, 6, 320, 0.1, "GREEDYNMM", "IOU", 0.3, 0.4),
]


def preds_to_json(preds):
    out = []
    for p in preds:
        k = getattr(p, "keypoints", None)
        out.append({"bbox": [int(v) for v in p.bbox.to_xyxy()], "score": float(p.score.value),
                    "category": [int(p.category.id), p.category.name],
                    "keypoints": None if k is None else np.asarray(k, dtype=np.float32).round(4).tolist()})
    return out


def main():
    predict, pred, base, yolo_wrapper, insight, enhancer = load_all()
    golden = {"yolo": {}, "insightface": {}, "enhancer": {}}
    for (name, H, W, nf, seed, sl, ov, ptype, metric, thr, conf) in SCENES:
        img = fd.coordinate_image(H, W)
        faces = fd.synthetic_faces(H, W, nf, seed)
        fd.FakeYOLO.faces = faces
        model = yolo_wrapper.YOLOv11PoseDetectionModel(model_path="fake.pt", confidence_threshold=conf, device="cpu", image_size=1024)
        res = predict.get_sliced_prediction(img, model, slice_height=sl, slice_width=sl, overlap_height_ratio=ov,
                                            overlap_width_ratio=ov, postprocess_type=ptype, postprocess_match_metric=metric,
                                            postprocess_match_threshold=thr, verbose=0)
        cache_keys = list(model.keypoints_cache.keys())
        out = model.attach_keypoints_to_predictions(res.object_prediction_list)
        golden["yolo"][name] = {"params": [H, W, nf, seed, sl, ov, ptype, metric, thr, conf], "stage1_keys": cache_keys,
                                "merged": preds_to_json(out), "image_wh": [res.image_width, res.image_height]}
        # evaluator settings on the same scene: NMS / IOS / class-agnostic (eval/eval_official_widerface.py:200-207)
        fd.FakeFaceAnalysis.faces = faces
        m2 = insight.InsightFaceDetectionModel(confidence_threshold=conf, providers=["CPUExecutionProvider"])
        res2 = predict.get_sliced_prediction(img, m2, slice_height=sl, slice_width=sl, overlap_height_ratio=ov,
                                             overlap_width_ratio=ov, postprocess_type="NMS", postprocess_match_metric="IOU",
                                             postprocess_match_threshold=0.5, postprocess_class_agnostic=True, verbose=0)
        golden["insightface"][name] = {"merged": preds_to_json(res2.object_prediction_list)}
    # FaceEnhancer (CPU branch of the reference: half off, tile capped at 200) over the exact affine up-sampler
    rng = np.random.default_rng(7)
    for name, (h, w), mname, scale, tile in [("x4_crop", (37, 53), "RealESRGAN_x4plus", 4, 256),
                                             ("x2_odd_tiled", (231, 317), "RealESRGAN_x2plus", 2, 400),
                                             ("x4_tiled", (210, 260), "RealESRGAN_x4plus", 4, 100)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        fe = enhancer.FaceEnhancer(model_name=mname, model_path="missing.pth", scale=scale, tile=tile, half=False)
        fe.upsampler.model_path = None
        out, ok = fe.enhance_image(img)
        golden["enhancer"][name] = {"seed_order": name, "shape": [h, w], "model": mname, "scale": fe.scale, "tile": fe.tile,
                                    "ok": bool(ok), "out_shape": list(out.shape), "sha_sum": int(out.astype(np.int64).sum()),
                                    "crc": int(np.bitwise_xor.reduce((out.astype(np.int64).ravel() * (np.arange(out.size) % 8191 + 1)) % 1000003))}
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote", os.path.join(HERE, "reference_outputs.json"),
          {k: len(v) for k, v in golden.items()})


if __name__ == "__main__":
    main()
