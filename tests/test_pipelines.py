"""(f3) pipeline glue: integer helpers against the oracle / hand-computed values (CPU) and the device-resident
enhancement-first / detection-first pipelines against the oracle flow (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import fake_detectors as fd  # noqa: E402

from oracle import pipelines as op  # noqa: E402


def test_slice_choosers_match_reference_rules():
    import fsd_b200.pipelines as pp

    assert pp.choose_slice_params(7680, 4320) == (1088, 1920, 0.2, 0.2)       # 4x4: ceil(7680/4)=1920, ceil(4320/4)=1080 -> 1088
    assert pp.choose_slice_params(2560, 1440) == (512, 896, 0.2, 0.2)         # 3x3: 854 -> 896, 480 -> 512
    assert pp.choose_slice_params(100, 50) == (50, 64, 0.2, 0.2)              # capped at the image
    for w, h, prefer in [(1920, 1080, "auto"), (3000, 2000, "auto"), (2999, 100, "4x4"), (5000, 5000, "3x3"), (63, 65, "auto")]:
        assert pp.choose_slice_params(w, h, prefer) == op.choose_slice_params(w, h, prefer)
    for w, h in [(1024, 768), (1501, 10), (2500, 2500), (2501, 1), (4160, 2340)]:
        assert pp.adaptive_slice_size(w, h) == op.adaptive_slice_size(w, h)
    assert [pp.adaptive_slice_size(1024, 768), pp.adaptive_slice_size(1920, 1080), pp.adaptive_slice_size(3840, 2160)] == [320, 416, 512]
    for d in (1, 2, 300, 767, 768, 2000):
        assert pp.detection_first_slice_size(d) == op.detection_first_slice_size(d)
    boxes = [[-5.5, 3.9, 20.2, 30.9], [90, 90, 150, 150], [10, 10, 10, 50], [40.7, 40.2, 41.9, 41.1]]
    assert pp.crop_rectangles(boxes, 100, 100) == op.crop_rectangles(boxes, 100, 100) == [(0, 3, 20, 30), (90, 90, 100, 100), (40, 40, 41, 41)]


@pytest.mark.gpu
def test_enhancement_first_pipeline_on_device(cuda_device):
    """Config 5 flow: x2 enhancement (exact affine up-sampler) -> sliced detection on the enhanced image -> boxes / 2,
    against the oracle flow (oracle RealESRGANer + oracle get_sliced_prediction over the same fake detector)."""
    import fsd_b200.pipelines as pp
    from fsd_b200.enhancer import RealESRGANer
    from fsd_b200.plugins import InsightFaceDetectionModel
    from fsd_b200.sahi_api import get_sliced_prediction
    from oracle import esrgan as oesr
    from oracle import predict as opred
    from oracle import yolo_wrapper as owrap

    H, W, scale = 270, 480, 2
    img = fd.coordinate_image(H, W)
    up = RealESRGANer(scale=scale, model=fd.AffineUpsampler(scale=scale), tile=128, tile_pad=10, pre_pad=0, half=False)
    big = up.enhance_device(torch.from_numpy(img).to(cuda_device)).cpu().numpy()
    want_big, _ = oesr.RealESRGANer(scale=scale, model=fd.AffineUpsampler(scale=scale), tile=128, tile_pad=10, pre_pad=0).enhance(img, outscale=scale)
    assert np.array_equal(big, want_big)
    # detection on the enhanced image through the generic plug-in path, then the reference's back-projection
    sh, sw, ovh, ovw = pp.choose_slice_params(W * scale, H * scale)
    fd.FakeFaceAnalysis.faces = fd.synthetic_faces(H * scale, W * scale, 20, 5)
    coord = fd.coordinate_image(H * scale, W * scale)  # the fake detector needs coordinate-coded pixels
    got = get_sliced_prediction(coord, InsightFaceDetectionModel(model=fd.FakeFaceAnalysis(), confidence_threshold=0.4),
                                slice_height=sh, slice_width=sw, overlap_height_ratio=ovh, overlap_width_ratio=ovw, verbose=0)
    want = opred.get_sliced_prediction(coord, owrap.InsightFaceDetectionModel(model=fd.FakeFaceAnalysis(), confidence_threshold=0.4),
                                       slice_height=sh, slice_width=sw, overlap_height_ratio=ovh, overlap_width_ratio=ovw, verbose=0)
    pp.rescale_boxes_(got.object_prediction_list, scale)
    assert [p.bbox.to_xyxy() for p in got.object_prediction_list] == [[v / scale for v in p.bbox.to_xyxy()] for p in want.object_prediction_list]
    assert len(got.object_prediction_list) > 5


@pytest.mark.gpu
def test_device_pipelines_run_end_to_end(cuda_device):
    """enhance_then_detect / detect_then_enhance with the real (random-init) networks: shapes, types, coordinate ranges."""
    import fsd_b200.pipelines as pp
    from fsd_b200.enhancer import FaceEnhancer
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.synthetic import make_image
    from fsd_b200.yolo import YOLO

    img, _ = make_image(11, 160, 224)
    model = YOLOv11PoseDetectionModel(model=YOLO("random-init"), confidence_threshold=0.4, device="cuda:0", image_size=512)
    fe = FaceEnhancer(model_name="RealESRGAN_x2plus", scale=2, tile=96, half=True, allow_random_init=True)
    res, big = pp.enhance_then_detect(img, fe, model, slice_params=(256, 256, 0.2, 0.2))
    assert tuple(big.shape) == (320, 448, 3) and (res.image_width, res.image_height) == (224, 160)
    for p in res.object_prediction_list:
        x1, y1, x2, y2 = p.bbox.to_xyxy()
        assert 0 <= x1 <= x2 <= 224 and 0 <= y1 <= y2 <= 160
    res2, crops = pp.detect_then_enhance(img, model, fe)
    rects = pp.crop_rectangles([p.bbox.to_xyxy() for p in res2.object_prediction_list], 224, 160)
    assert len(crops) == len(rects)
    for (x1, y1, x2, y2), c in zip(rects, crops):
        s = 2 if (y2 - y1 >= 4 and x2 - x1 >= 4) else 1
        assert c.shape == ((y2 - y1) * s, (x2 - x1) * s, 3) and c.dtype == np.uint8
