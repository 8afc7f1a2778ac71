"""End-to-end parity of the fused device path against the reference-equivalent CPU flow, with the backbone factored out:
the GPU pipeline records the head tensors of every network input; the oracle replays them (its own letterbox output must
equal Kernel 1's, bit for bit, to find them) through ultralytics-style decode/NMS/rescale, the plug-in's int()/shift,
sahi's merge and the plug-in's key-point attach.  Boxes, groupings and WIDER-FACE-style AP must then be identical."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

CASES = [  # H, W, slice, overlap, imgsz, conf, postprocess, metric
    (384, 512, 256, 0.2, 512, 0.5, "GREEDYNMM", "IOS"),
    (384, 512, 256, 0.2, 512, 0.05, "NMS", "IOS"),        # evaluator settings: low confidence -> max_det pressure
    (300, 500, 320, 0.25, 640, 0.3, "NMM", "IOU"),         # non-square letterbox padding, NMM
    (300, 500, 200, 0.2, 320, 0.4, "GREEDYNMM", "IOU"),    # 1.6x slices (SAHI's 640 -> 1024 ratio): Kernel 1's sixteenths path
]


from parity_utils import Recorder, as_rows, run_sliced_case  # noqa: E402


@pytest.mark.parametrize("case", CASES)
def test_fused_path_equals_oracle_flow(cuda_device, case):
    """Small geometries (every merge type / metric); the BASELINE configurations at full size are in
    tests/test_parity_fullsize_gpu.py.  No image is skipped: see parity_utils for the int()-flip rule."""
    H, W, sl, ov, imgsz, conf, ptype, metric = case
    out = run_sliced_case(H, W, sl, ov, imgsz, conf, ptype, metric, n_images=4)
    assert out["boxes"] > 20 and out["flip_images"] <= 1, out


def test_batch_api_equals_single_image_api(cuda_device):
    from fsd_b200.api import get_sliced_prediction_batch
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.sahi_api import get_sliced_prediction
    from fsd_b200.synthetic import make_image
    from fsd_b200.yolo import YOLO

    model = YOLOv11PoseDetectionModel(model=YOLO("random-init"), confidence_threshold=0.4, device="cuda:0", image_size=512)
    imgs = [make_image(200 + i, 384, 512)[0] for i in range(5)]
    batch = get_sliced_prediction_batch(imgs, model, 256, 256, 0.2, 0.2)
    # a different batch size changes cuDNN's algorithm choice, hence (slightly) the head tensors: compare loosely here,
    # the exact comparison is test_fused_path_equals_oracle_flow
    n_kp = n_all = 0
    for img, res in zip(imgs, batch):
        one = get_sliced_prediction(img, model, slice_height=256, slice_width=256, verbose=0)
        assert abs(len(one.object_prediction_list) - len(res.object_prediction_list)) <= max(5, len(one.object_prediction_list) // 3)
        assert (res.image_width, res.image_height) == (512, 384)
        n_kp += sum(hasattr(p, "keypoints") for p in res.object_prediction_list)  # union boxes may match no detection
        n_all += len(res.object_prediction_list)
    assert n_all > 0 and n_kp >= n_all // 2


@pytest.mark.parametrize("use_graphs,overlap_post,depth", [(True, True, 3), (False, True, 2), (True, False, 2), (False, False, 1)])
def test_predict_stream_equals_batch_api(cuda_device, use_graphs, overlap_post, depth):
    """The pipelined API (2 batches in flight on separate streams; backbone chunks replayed as CUDA graphs over static
    per-stream buffers, or launched eagerly) returns exactly what the synchronous batch call returns."""
    from fsd_b200.api import get_sliced_prediction_batch, predict_stream
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.synthetic import make_image
    from fsd_b200.yolo import YOLO

    model = YOLOv11PoseDetectionModel(model=YOLO("random-init"), confidence_threshold=0.4, device="cuda:0", image_size=512)
    imgs = [torch.from_numpy(make_image(300 + i, 384, 512)[0]).pin_memory() for i in range(12)]
    groups = [imgs[0:4], imgs[4:8], imgs[8:12], imgs[0:4], imgs[4:8], imgs[8:12]]
    want = [get_sliced_prediction_batch(g, model, 256, 256, 0.2, 0.2) for g in groups]
    # the synchronous call with its slices' stage-1 work on a side stream (engine.overlap_post) returns the same objects
    model.engine().overlap_post = True
    again = get_sliced_prediction_batch(groups[1], model, 256, 256, 0.2, 0.2)
    model.engine().overlap_post = False
    for gr, wr in zip(again, want[1]):
        assert [(p.bbox.to_xyxy(), p.score.value) for p in gr.object_prediction_list] == [(p.bbox.to_xyxy(), p.score.value) for p in wr.object_prediction_list]
    stats = {}
    got = list(predict_stream(iter(groups), model, 256, 256, 0.2, 0.2, depth=depth, rows_per_image_hint=8,  # tiny window: exercises the refetch
                              stats=stats, use_graphs=use_graphs, overlap_post=overlap_post))
    eng = model.engine()
    assert eng.use_graphs is False and (eng.replayed_launches > 0) == use_graphs or not use_graphs
    # batches that already live on the device (ImagePool) go through the same pipeline without the upload
    import fsd_b200.ops as ops

    pools = [ops.ImagePool.from_numpy([im.numpy() for im in g], cuda_device) for g in groups[:3]]
    res = list(predict_stream(iter(pools), model, 256, 256, 0.2, 0.2, depth=depth, use_graphs=use_graphs, overlap_post=overlap_post))
    for gb, wb in zip(res, want[:3]):
        for gr, wr in zip(gb, wb):
            assert [p.bbox.to_xyxy() for p in gr.object_prediction_list] == [p.bbox.to_xyxy() for p in wr.object_prediction_list]
            assert (gr.image_width, gr.image_height) == (512, 384)
    assert set(stats) == {"enqueue", "wait", "build", "device", "device_idle"} and all(v >= 0 for v in stats.values())
    assert len(got) == len(want)
    for gb, wb in zip(got, want):
        for gr, wr in zip(gb, wb):
            a = [(p.bbox.to_xyxy(), round(p.score.value, 6), hasattr(p, "keypoints")) for p in gr.object_prediction_list]
            b = [(p.bbox.to_xyxy(), round(p.score.value, 6), hasattr(p, "keypoints")) for p in wr.object_prediction_list]
            assert a == b


def test_fused_path_edge_cases(cuda_device):
    """No detections at all; an image smaller than the slice (one slice, no full-image pass); image with an odd row pitch."""
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.sahi_api import get_sliced_prediction
    from fsd_b200.synthetic import make_image
    from fsd_b200.yolo import YOLO
    from oracle import predict as opred
    from oracle.yolo_head import OracleYOLO
    from oracle.yolo_wrapper import YOLOv11PoseDetectionModel as OracleModel

    yolo = YOLO("random-init")
    # (1) confidence so high that nothing survives: empty list, no error, image size still reported
    model = YOLOv11PoseDetectionModel(model=yolo, confidence_threshold=1.0, device="cuda:0", image_size=512)
    img, _ = make_image(7, 300, 401)  # 401*3 = 1203 bytes per row: not a multiple of 16 -> pitched pool
    res = get_sliced_prediction(img, model, slice_height=256, slice_width=256, verbose=0)
    assert res.object_prediction_list == [] and (res.image_width, res.image_height) == (401, 300)
    assert model.attach_keypoints_to_predictions(res.object_prediction_list) == []
    # (2) image smaller than the slice: exactly one slice and no standard pass, compared with the oracle flow
    model = YOLOv11PoseDetectionModel(model=yolo, confidence_threshold=0.3, device="cuda:0", image_size=512)
    eng = model.engine()
    rec = Recorder()
    eng.head_hook = rec.hook
    got = get_sliced_prediction(img, model, slice_height=512, slice_width=512, verbose=0)
    eng.head_hook = None
    assert len(rec.items) == 1
    omodel = OracleModel(model=OracleYOLO(None, half=True, head_hook=rec.lookup), confidence_threshold=0.3, device="cpu", image_size=512)
    want = opred.get_sliced_prediction(img, omodel, slice_height=512, slice_width=512, verbose=0)
    assert [r[0] for r in as_rows(got.object_prediction_list)] == [r[0] for r in as_rows(want.object_prediction_list)]
    assert len(got.object_prediction_list) > 0


@pytest.mark.parametrize("cfg", [("C2", 768, 1024, 512, 4), ("C1", 1080, 1920, 640, 2)], ids=lambda c: c[0])
def test_full_size_properties(cuda_device, cfg):
    """BASELINE configs at their FULL sizes (imgsz 1024), through properties that do not need the CPU oracle to finish:
    determinism (two runs, and graph replay vs eager launches, give identical rows), geometric sanity of every box,
    stage-1 caps, and idempotence of the merge (merging the merged NMS output again keeps everything)."""
    import fsd_b200.ops as ops
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.synthetic import make_image
    from fsd_b200.yolo import YOLO

    name, H, W, sl, n = cfg
    torch.backends.cudnn.deterministic = True
    model = YOLOv11PoseDetectionModel(model=YOLO("random-init"), confidence_threshold=0.5, device="cuda:0", image_size=1024)
    eng = model.engine()
    eng.truncate = True  # the plug-in's int() truncation (utils/yolo_wrapper.py:138), as the reference-facing API sets it
    pool = ops.ImagePool.from_numpy([make_image(900 + i, H, W, mean_faces=12)[0] for i in range(n)], cuda_device)
    kw = dict(postprocess_type="NMS", match_metric="IOS", match_threshold=0.5)
    a = eng.detect(pool, sl, sl, 0.2, 0.2, True, want_stage1=True, **kw)
    b = eng.detect(pool, sl, sl, 0.2, 0.2, True, **kw)
    eng.use_graphs = True
    try:
        c = eng.detect(pool, sl, sl, 0.2, 0.2, True, **kw)   # first call captures the graphs, second replays them
        d = eng.detect(pool, sl, sl, 0.2, 0.2, True, **kw)
    finally:
        eng.use_graphs = False
        torch.backends.cudnn.deterministic = False
    for other in (b, c, d):
        assert np.array_equal(a.offsets, other.offsets) and np.array_equal(a.boxes, other.boxes) and np.array_equal(a.scores, other.scores)
        assert np.array_equal(a.keypoints, other.keypoints)
    S = a.counters["slices"]
    assert S == {"C2": 6, "C1": 8}[name] and a.counters["entries"] == n * (S + 1)
    assert int(a.offsets[-1]) > 0
    bx = a.boxes
    assert (bx == np.floor(bx)).all() and (bx[:, 0] >= 0).all() and (bx[:, 1] >= 0).all()
    assert (bx[:, 2] <= W).all() and (bx[:, 3] <= H).all() and (bx[:, 2] > bx[:, 0]).all() and (bx[:, 3] > bx[:, 1]).all()
    assert (a.scores >= 0.5).all() and (a.scores <= 1.0).all()
    assert all(int(cnt) <= (S + 1) * 300 for cnt in a.stage1["count"])
    # idempotence: the NMS/IOS output of every image, merged again with the same rule, keeps every box
    for i in range(n):
        lo, hi = int(a.offsets[i]), int(a.offsets[i + 1])
        if hi - lo < 2:
            continue
        rows = torch.zeros((hi - lo, 6), device=cuda_device)
        rows[:, :4] = torch.from_numpy(bx[lo:hi]).to(cuda_device)
        rows[:, 4] = torch.from_numpy(a.scores[lo:hi]).to(cuda_device)
        res = ops.merge_segments(rows, torch.zeros(1, dtype=torch.int32, device=cuda_device), None, hi - lo, merge_type="NMS",
                                 metric="IOS", thr=0.5, precision="fp64", tie_rule=eng.tie_rule)
        assert int(res["keep_count"][0]) == hi - lo
