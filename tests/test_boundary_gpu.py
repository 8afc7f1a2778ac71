"""Boundary tests on the GPU: the product's reference-facing API must give the outputs recorded from the reference's own
files (tests/golden/reference_outputs.json) for the same plug-in detections, with shift + merge + key-point attach running
through Kernel 3 / the attach kernel instead of sahi's CPU code."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import fake_detectors as fd  # noqa: E402

pytestmark = pytest.mark.gpu
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
def test_yolo_plugin_generic_path_matches_reference(cuda_device, name):
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.sahi_api import get_sliced_prediction

    g = GOLD["yolo"][name]
    H, W, nf, seed, sl, ov, ptype, metric, thr, conf = g["params"]
    img = fd.coordinate_image(H, W)
    fd.FakeYOLO.faces = fd.synthetic_faces(H, W, nf, seed)
    model = YOLOv11PoseDetectionModel(model=fd.FakeYOLO(), confidence_threshold=conf, device="cuda:0", image_size=1024)
    assert not model.supports_batched_slices
    res = get_sliced_prediction(img, model, slice_height=sl, slice_width=sl, overlap_height_ratio=ov,
                                overlap_width_ratio=ov, postprocess_type=ptype, postprocess_match_metric=metric,
                                postprocess_match_threshold=thr, verbose=0)
    assert list(model.keypoints_cache.keys()) == g["stage1_keys"]
    out = model.attach_keypoints_to_predictions(res.object_prediction_list)
    assert as_json(out) == g["merged"]
    assert [res.image_width, res.image_height] == g["image_wh"]
    assert set(res.durations_in_seconds) == {"slice", "prediction", "postprocess"}


@pytest.mark.parametrize("name", sorted(GOLD["insightface"]))
def test_insightface_plugin_matches_reference(cuda_device, name):
    from fsd_b200.plugins import InsightFaceDetectionModel
    from fsd_b200.sahi_api import get_sliced_prediction

    H, W, nf, seed, sl, ov, _, _, _, conf = GOLD["yolo"][name]["params"]
    img = fd.coordinate_image(H, W)
    fd.FakeFaceAnalysis.faces = fd.synthetic_faces(H, W, nf, seed)
    model = InsightFaceDetectionModel(model=fd.FakeFaceAnalysis(), confidence_threshold=conf)
    res = get_sliced_prediction(img, model, slice_height=sl, slice_width=sl, overlap_height_ratio=ov,
                                overlap_width_ratio=ov, postprocess_type="NMS", postprocess_match_metric="IOU",
                                postprocess_match_threshold=0.5, postprocess_class_agnostic=True, verbose=0)
    assert as_json(res.object_prediction_list) == GOLD["insightface"][name]["merged"]


def test_merge_buffer_length_and_lsnms(cuda_device):
    from fsd_b200.plugins import InsightFaceDetectionModel
    from fsd_b200.sahi_api import get_sliced_prediction
    from oracle import predict as opred
    from oracle import yolo_wrapper as owrap

    img = fd.coordinate_image(480, 640)
    fd.FakeFaceAnalysis.faces = fd.synthetic_faces(480, 640, 30, 11)
    kw = dict(slice_height=256, slice_width=256, overlap_height_ratio=0.2, overlap_width_ratio=0.2, merge_buffer_length=8, verbose=0)
    got = get_sliced_prediction(img, InsightFaceDetectionModel(model=fd.FakeFaceAnalysis(), confidence_threshold=0.4), **kw)
    want = opred.get_sliced_prediction(img, owrap.InsightFaceDetectionModel(model=fd.FakeFaceAnalysis(), confidence_threshold=0.4), **kw)
    assert as_json(got.object_prediction_list) == as_json(want.object_prediction_list)
    with pytest.raises(NotImplementedError):
        get_sliced_prediction(img, InsightFaceDetectionModel(model=fd.FakeFaceAnalysis(), confidence_threshold=0.4),
                              slice_height=256, slice_width=256, postprocess_type="LSNMS", verbose=0)


def enhancer_cases():
    rng = np.random.default_rng(7)
    for name, (h, w), scale, tile in [("x4_crop", (37, 53), 4, 200), ("x2_odd_tiled", (231, 317), 2, 200), ("x4_tiled", (210, 260), 4, 100)]:
        yield name, rng.integers(0, 256, (h, w, 3), dtype=np.uint8), scale, tile


def test_enhancer_matches_reference(cuda_device):
    """RealESRGANer on the GPU (Kernel 4 crop/stitch, tiles batched per shape) == the reference's FaceEnhancer output."""
    from fsd_b200.enhancer import RealESRGANer

    for name, img, scale, tile in enhancer_cases():
        g = GOLD["enhancer"][name]
        up = RealESRGANer(scale=scale, model=fd.AffineUpsampler(scale=scale), tile=tile, tile_pad=10, pre_pad=0, half=False)
        out, mode = up.enhance(img, outscale=scale)
        crc = int(np.bitwise_xor.reduce((out.astype(np.int64).ravel() * (np.arange(out.size) % 8191 + 1)) % 1000003))
        assert mode == "RGB" and [list(out.shape), int(out.astype(np.int64).sum()), crc] == [g["out_shape"], g["sha_sum"], g["crc"]]
        up.tile = 50  # the reference's OOM retry assigns `.tile`; it must stay assignable (and stay a no-op)
        assert up.tile_size == tile


def test_face_enhancer_contract(cuda_device):
    from fsd_b200.enhancer import FaceEnhancer

    fe = FaceEnhancer(model_name="RealESRGAN_x2plus", model_path=None, scale=4, tile=64, half=True, allow_random_init=True)
    assert fe.scale == 2 and fe.device == "cuda" and fe.get_model_info()["is_loaded"]
    img = np.random.default_rng(0).integers(0, 256, (70, 90, 3), dtype=np.uint8)
    out, ok = fe.enhance_image(img)
    assert ok and out.shape == (140, 180, 3) and out.dtype == np.uint8
    tiny, ok = fe.enhance_image(img[:3, :3])
    assert not ok and tiny.shape == (3, 3, 3)  # (input, False) instead of raising
    from PIL import Image

    out2, ok = fe.enhance_image(Image.fromarray(img[:, :, ::-1].copy()))  # PIL input is converted RGB -> BGR first
    assert ok and np.array_equal(out2, out)
