"""The oracle must reproduce the outputs of the reference's OWN files (docs sahi/{predict,prediction,base}.py,
utils/{yolo_wrapper,insightface_wrapper,enhancer}.py), recorded by tests/golden/make_golden.py which imports them
unmodified from /root/reference.  This is the pin that makes the oracle trustworthy for the in-repo half of the path."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import fake_detectors as fd  # noqa: E402

from oracle import esrgan as oesr  # noqa: E402
from oracle import predict as opred  # noqa: E402
from oracle import yolo_wrapper as owrap  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "reference_outputs.json")))


def as_json(preds):
    out = []
    for p in preds:
        k = getattr(p, "keypoints", None)
        out.append({"bbox": [int(v) for v in p.bbox.to_xyxy()], "score": float(p.score.value),
                    "category": [int(p.category.id), p.category.name],
                    "keypoints": None if k is None else np.asarray(k, dtype=np.float32).round(4).tolist()})
    return out


@pytest.mark.parametrize("name", sorted(GOLD["yolo"]))
def test_yolo_plugin_sliced_prediction_matches_reference(name):
    g = GOLD["yolo"][name]
    H, W, nf, seed, sl, ov, ptype, metric, thr, conf = g["params"]
    img = fd.coordinate_image(H, W)
    fd.FakeYOLO.faces = fd.synthetic_faces(H, W, nf, seed)
    model = owrap.YOLOv11PoseDetectionModel(model=fd.FakeYOLO(), confidence_threshold=conf, device="cpu", image_size=1024)
    res = opred.get_sliced_prediction(img, model, slice_height=sl, slice_width=sl, overlap_height_ratio=ov,
                                      overlap_width_ratio=ov, postprocess_type=ptype, postprocess_match_metric=metric,
                                      postprocess_match_threshold=thr, verbose=0)
    assert list(model.keypoints_cache.keys()) == g["stage1_keys"]
    out = model.attach_keypoints_to_predictions(res.object_prediction_list)
    assert as_json(out) == g["merged"]
    assert [res.image_width, res.image_height] == g["image_wh"]


@pytest.mark.parametrize("name", sorted(GOLD["insightface"]))
def test_insightface_plugin_matches_reference(name):
    H, W, nf, seed, sl, ov, _, _, _, conf = GOLD["yolo"][name]["params"]
    img = fd.coordinate_image(H, W)
    fd.FakeFaceAnalysis.faces = fd.synthetic_faces(H, W, nf, seed)
    fa = fd.FakeFaceAnalysis()
    model = owrap.InsightFaceDetectionModel(model=fa, confidence_threshold=conf)
    res = opred.get_sliced_prediction(img, model, slice_height=sl, slice_width=sl, overlap_height_ratio=ov,
                                      overlap_width_ratio=ov, postprocess_type="NMS", postprocess_match_metric="IOU",
                                      postprocess_match_threshold=0.5, postprocess_class_agnostic=True, verbose=0)
    assert as_json(res.object_prediction_list) == GOLD["insightface"][name]["merged"]


def enhancer_cases():
    rng = np.random.default_rng(7)
    for name, (h, w), scale, tile in [("x4_crop", (37, 53), 4, 200), ("x2_odd_tiled", (231, 317), 2, 200), ("x4_tiled", (210, 260), 4, 100)]:
        yield name, rng.integers(0, 256, (h, w, 3), dtype=np.uint8), scale, tile


def checksum(out):
    return [list(out.shape), int(out.astype(np.int64).sum()),
            int(np.bitwise_xor.reduce((out.astype(np.int64).ravel() * (np.arange(out.size) % 8191 + 1)) % 1000003))]


def test_enhancer_matches_reference():
    """utils/enhancer.py FaceEnhancer on its CPU branch (half off, tile <= 200, tile_pad 10, pre_pad 0)."""
    for name, img, scale, tile in enhancer_cases():
        g = GOLD["enhancer"][name]
        assert (g["scale"], g["tile"]) == (scale, tile)
        up = oesr.RealESRGANer(scale=scale, model=fd.AffineUpsampler(scale=scale), tile=tile, tile_pad=10, pre_pad=0, half=False)
        out, _ = up.enhance(img, outscale=scale)
        assert checksum(out) == [g["out_shape"], g["sha_sum"], g["crc"]]
