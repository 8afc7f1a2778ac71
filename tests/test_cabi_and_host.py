"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol include/fsd_b200.h
declares; host-side planners work without a GPU; the product fails LOUDLY (no CPU fallback) when no device exists;
the mirrored SAHI data classes behave like the oracle's (= the reference's) classes."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "fsd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import fsd_b200._cabi as cabi

    lib = cabi.load_library()
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/fsd_b200.h but not exported by libfsd_b200.so"
        assert n in cabi.SIGNATURES, f"{n} has no ctypes prototype"
    assert cabi.MISSING == []
    assert sorted(cabi.SIGNATURES) == names, "ctypes table and header drifted apart"
    assert b"sm_100a" in lib.fsd_version()


def test_library_is_sm100a_only():
    import subprocess

    import fsd_b200._cabi as cabi

    out = subprocess.run(["cuobjdump", "-lelf", str(cabi.library_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import fsd_b200._cabi as cabi
    from fsd_b200.sahi_api import NMSPostprocess, ObjectPrediction

    lib = cabi.load_library()
    h = ctypes.c_void_p()
    rc = lib.fsd_create(0, ctypes.byref(h))
    assert rc == -5 and b"no CPU fallback" in lib.fsd_last_error()
    with pytest.raises(cabi.FsdError):
        cabi.Handle(0)
    preds = [ObjectPrediction(bbox=[0, 0, 5, 5], category_id=0, score=0.5), ObjectPrediction(bbox=[1, 1, 6, 6], category_id=0, score=0.4)]
    with pytest.raises(cabi.FsdError):
        NMSPostprocess()(preds)
    from fsd_b200.yolo import YOLO

    with pytest.raises(RuntimeError):
        YOLO("random-init").predict(np.zeros((32, 32, 3), np.uint8))


def test_argument_errors_are_reported_without_a_gpu():
    import fsd_b200._cabi as cabi

    lib = cabi.load_library()
    n = ctypes.c_int(0)
    assert lib.fsd_slice_plan(0, 10, 4, 4, 0.2, 0.2, None, 0, ctypes.byref(n)) == -1
    assert b"sizes must be" in lib.fsd_last_error()
    buf = (ctypes.c_int32 * 4)()
    assert lib.fsd_slice_plan(100, 100, 40, 40, 0.2, 0.2, buf, 1, ctypes.byref(n)) == -4 and n.value == 9
    assert lib.fsd_merge_workspace_bytes(100, 2, 5000) == 2 * (8192 * 52 + 2048) + 256  # box arrays (+ NMM step array) + the cluster kernel's scratch
    tab = (ctypes.c_int32 * 12)()
    assert lib.fsd_esrgan_tile_table(3, 3, 2, 2, 10, 5, tab, 1, ctypes.byref(n), None) == -1  # reflect pad >= size


def test_mirrored_data_classes_match_oracle():
    from fsd_b200.sahi_api import BoundingBox, ObjectPrediction, PredictionResult, PredictionScore
    from oracle import annotation as oa

    for cls_p, cls_o in ((ObjectPrediction, oa.ObjectPrediction),):
        for bbox, shift, full in ([[-3, 5, 700, 90], [10, 20], [80, 640]], [np.array([1, 2, 30, 40]), [0, 0], None],
                                  [[5.5, 6.5, 7.5, 8.5], [100, 200], [1000, 1000]]):
            p = cls_p(bbox=bbox, category_id=0, category_name="face", score=np.float32(0.75), shift_amount=shift, full_shape=full)
            o = cls_o(bbox=bbox, category_id=0, category_name="face", score=np.float32(0.75), shift_amount=shift, full_shape=full)
            assert p.bbox.to_xyxy() == o.bbox.to_xyxy() and p.bbox.to_xywh() == o.bbox.to_xywh()
            assert p.get_shifted_object_prediction().bbox.to_xyxy() == o.get_shifted_object_prediction().bbox.to_xyxy()
            assert p.score.value == o.score.value and isinstance(p.score.value, float)
            assert (p.category.id, p.category.name) == (o.category.id, o.category.name)
            assert p.get_shifted_object_prediction().bbox.shift_amount == (0, 0)
    with pytest.raises(Exception):
        BoundingBox([1, 2, -3, 4])
    s = PredictionScore(np.float32(0.5))
    assert s > 0.4 and s < 0.6 and s == 0.5 and s.is_greater_than_threshold(0.1)
    r = PredictionResult([], np.zeros((12, 34, 3), np.uint8), {"slice": 0})
    assert (r.image_width, r.image_height) == (34, 12) and r.image.size == (34, 12)


def test_detection_model_protocol():
    from fsd_b200.sahi_api import DetectionModel, ObjectPrediction

    class Dummy(DetectionModel):
        def load_model(self):
            self.model = "m"

        def perform_inference(self, image):
            self._original_predictions = [image.shape]

        def _create_object_prediction_list_from_original_predictions(self, shift_amount_list=[[0, 0]], full_shape_list=None):
            self._object_prediction_list_per_image = [[ObjectPrediction(bbox=[1, 2, 3, 4], category_id=0, score=0.9,
                                                                        shift_amount=shift_amount_list, full_shape=full_shape_list)]]

    m = Dummy(confidence_threshold=0.4, device="cpu", category_remapping={"0": 7})
    assert m.model == "m" and m.object_prediction_list == [] and m.object_prediction_list_per_image == []
    m.perform_inference(np.zeros((4, 4, 3), np.uint8))
    m.convert_original_predictions(shift_amount=[5, 6], full_shape=[100, 100])
    op = m.object_prediction_list[0]
    assert op.category.id == 7 and op.get_shifted_object_prediction().bbox.to_xyxy() == [6, 8, 8, 10]
    assert m.original_predictions == [(4, 4, 3)]


def test_get_sliced_prediction_argument_errors():
    from fsd_b200.sahi_api import get_sliced_prediction

    with pytest.raises(ValueError, match="postprocess_type should be one of"):
        get_sliced_prediction(np.zeros((8, 8, 3), np.uint8), None, 4, 4, postprocess_type="FOO")


def test_convolution_shape_rules_and_tap_major_layouts():
    """The library owns the shape rules of the tensor-core convolutions (no GPU needed); the tap-major weight helpers are pure layout."""
    import torch

    import fsd_b200._cabi as cabi
    import fsd_b200.ops as ops

    lib = cabi.load_library()
    assert lib.fsd_pointwise_conv_supported(64, 64) == 2 and lib.fsd_pointwise_conv_supported(384, 128) == 2
    assert lib.fsd_pointwise_conv_supported(512, 256) == 0 and lib.fsd_pointwise_conv_supported(24, 64) == 0  # 256 KB of weights; K % 16
    assert lib.fsd_pointwise_conv_supported(128, 256) == 2 and lib.fsd_pointwise_conv_supported(256, 256) == 0
    assert lib.fsd_conv3x3_supported(64, 64) == 1 and lib.fsd_conv3x3_supported(16, 8) == 1 and lib.fsd_conv3x3_supported(8, 16) == 0
    assert lib.fsd_conv3x3_supported(128, 128) == 0 and lib.fsd_conv3x3_supported(256, 16) == 1
    assert lib.fsd_conv2x2_supported(64, 32) == 1 and lib.fsd_conv2x2_supported(128, 32) == 0
    assert ops.conv3x3_preferred(64, 64) and not ops.conv3x3_preferred(128, 16)
    w = torch.arange(2 * 16 * 9, dtype=torch.float16).reshape(2, 16, 3, 3)
    t = ops.conv3x3_tap_major(torch.cat([w] * 8))  # N = 16
    assert t.shape == (3, 3, 16, 16) and torch.equal(t[1, 2, 0], w[0, :, 1, 2])
    t8 = ops.conv3x3_tap_major(torch.cat([w] * 4))  # N = 8 -> zero-padded to 16 rows
    assert t8.shape == (3, 3, 16, 16) and float(t8[:, :, 8:].abs().max()) == 0 and torch.equal(t8[0, 0, 1], w[1, :, 0, 0])
    dw = torch.arange(24 * 9, dtype=torch.float16).reshape(24, 1, 3, 3)
    td = ops.dwconv3x3_tap_major(dw)
    assert td.shape == (9, 24) and torch.equal(td[5], dw[:, 0, 1, 2])


def test_new_convolution_entry_points_validate_arguments_without_a_gpu():
    """fsd_conv3x3 / fsd_conv2x2 / fsd_dwconv3x3 reject bad arguments before touching the device (status -1 and a message)."""
    import fsd_b200._cabi as cabi

    lib = cabi.load_library()
    assert lib.fsd_conv3x3(None, None, 0, 1, 8, 8, None, None, None, 0, None, 0, 16, 16, 1, 0.0, cabi.FSD_F16, None) == -1
    assert b"fsd_conv3x3" in lib.fsd_last_error()
    assert lib.fsd_conv2x2(None, None, 0, 1, 8, 8, None, None, None, 0, 64, 32, 1, 0.0, cabi.FSD_F16, None) == -1
    assert b"fsd_conv2x2" in lib.fsd_last_error()
    assert lib.fsd_dwconv3x3(None, None, 0, 1, 8, 8, None, None, None, 0, 64, 1, 0.0, cabi.FSD_F16, None) == -1
    assert b"fsd_dwconv3x3" in lib.fsd_last_error()
