"""`docs sahi/retinaface_sahi.py` (BASELINE config 3's named file): the oracle restatement (CPU) and the product
mirror (GPU for the sliced call: its merge is Kernel 3) must reproduce what the reference's OWN class returns — quirks
included — as recorded by tests/golden/make_golden_retinaface.py."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import fake_detectors as fd  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "retinaface_sahi_outputs.json")))


def ops_json(preds):
    return [{"bbox": [int(v) for v in p.bbox.to_xyxy()], "shift": [int(v) for v in p.bbox.shift_amount],
             "score": float(p.score.value), "category": [int(p.category.id), p.category.name],
             "shifted": [int(v) for v in p.get_shifted_object_prediction().bbox.to_xyxy()]} for p in preds]


def _check_class(make, sliced=None):
    for c in GOLD["ctor"]:
        m = make(confidence_threshold=0.45, device=c["device"], ctx_id=c["ctx_id_arg"], image_size=c["image_size_arg"])
        det = m._resolved_det_size() if hasattr(m, "_resolved_det_size") else m.det_size()
        assert (m.ctx_id, m.image_size, list(det)) == (c["ctx_id"], c["image_size"], c["det_size"])
        assert (m.category_names, m.has_mask, m.model_name, list(m.original_predictions)) == (c["names"], c["has_mask"], c["model_name"], [])
    for name, g in GOLD["scenes"].items():
        H, W, nf, seed, sl, conf = g["params"]
        img = fd.coordinate_image(H, W)
        fd.FakeFaceAnalysis.faces = fd.synthetic_faces(H, W, nf, seed)
        m = make(confidence_threshold=conf, device="cpu", image_size=640)
        window = np.ascontiguousarray(img[64:64 + sl, 128:128 + sl])
        assert ops_json(m.perform_inference(window)) == g["direct"]
        assert len(m.original_predictions) == g["n_raw"]
        assert ops_json(m._create_object_prediction_list_from_original_predictions([128, 64], [H, W])) == g["created"]
        assert ops_json(m._create_object_prediction_list_from_original_predictions(None, None)) == g["created_default"]
        m.convert_original_predictions(shift_amount=[128, 64], full_shape=[H, W])
        assert len(m.object_prediction_list) == g["after_convert"] == 0   # quirk 1: nothing is stored
        assert any(r["shifted"] != r["bbox"] for r in g["created"])       # quirk 2: shifting again moves the box
        assert ops_json(m.perform_inference(window.astype(np.float32) / 255.0)) == g["float_input"]
        assert len(m.perform_inference(np.zeros((8, 8), np.uint8))) == g["bad_shape"] == 0
        assert len(m.perform_inference(np.zeros((0, 0, 3), np.uint8))) == g["empty"] == 0
        if sliced is not None:
            res = sliced(img, m, slice_height=sl, slice_width=sl, overlap_height_ratio=0.2, overlap_width_ratio=0.2,
                         postprocess_type="NMS", postprocess_match_metric="IOU", postprocess_match_threshold=0.5, verbose=0)
            assert ops_json(res.object_prediction_list) == g["sliced"] == []


def test_oracle_reproduces_reference_class():
    from oracle import predict as opred
    from oracle.retinaface_sahi import RetinaFaceSAHI

    _check_class(lambda **kw: RetinaFaceSAHI(model=fd.FakeFaceAnalysis(), **kw), opred.get_sliced_prediction)


def test_product_mirror_reproduces_reference_class_on_the_host_side():
    from fsd_b200.retinaface_sahi import RetinaFaceSAHI

    _check_class(lambda **kw: RetinaFaceSAHI(model=fd.FakeFaceAnalysis(), **kw))
    with pytest.raises(ImportError, match="insightface"):
        RetinaFaceSAHI()


@pytest.mark.gpu
def test_product_mirror_through_get_sliced_prediction(cuda_device):
    """The sliced call with the mirror: quirks on -> no detections, exactly like the reference; quirks off -> the working
    behaviour, identical to the InsightFace wrapper (utils/insightface_wrapper.py) and to the oracle flow."""
    from fsd_b200.plugins import InsightFaceDetectionModel
    from fsd_b200.retinaface_sahi import RetinaFaceSAHI
    from fsd_b200.sahi_api import get_sliced_prediction
    from oracle import predict as opred
    from oracle import yolo_wrapper as owrap

    _check_class(lambda **kw: RetinaFaceSAHI(model=fd.FakeFaceAnalysis(), **kw), get_sliced_prediction)
    H, W, nf, seed, sl, conf = GOLD["scenes"]["crowd_a"]["params"]
    img = fd.coordinate_image(H, W)
    fd.FakeFaceAnalysis.faces = fd.synthetic_faces(H, W, nf, seed)
    kw = dict(slice_height=sl, slice_width=sl, overlap_height_ratio=0.2, overlap_width_ratio=0.2, postprocess_type="NMS",
              postprocess_match_metric="IOU", postprocess_match_threshold=0.5, verbose=0)
    fixed = get_sliced_prediction(img, RetinaFaceSAHI(model=fd.FakeFaceAnalysis(), confidence_threshold=conf, reference_quirks=False), **kw)
    twin = get_sliced_prediction(img, InsightFaceDetectionModel(model=fd.FakeFaceAnalysis(), confidence_threshold=conf), **kw)
    want = opred.get_sliced_prediction(img, owrap.InsightFaceDetectionModel(model=fd.FakeFaceAnalysis(), confidence_threshold=conf), **kw)
    rows = lambda r: [([int(v) for v in p.bbox.to_xyxy()], float(p.score.value)) for p in r.object_prediction_list]  # noqa: E731
    assert rows(fixed) == rows(twin) == rows(want) and len(rows(fixed)) > 10
