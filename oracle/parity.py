"""Oracle-side parity machinery (TEST INFRASTRUCTURE, see oracle/__init__): the fused device path against the
reference-equivalent CPU flow.  Used by tests/ (through tests/parity_utils.py) and by bench.py's parity gate — as the
checker only.

The backbone is factored out: the GPU pipeline records the head tensors of every network input (`Recorder.hook`), the
oracle replays them (`Recorder.lookup` finds them by the bit-identical letterboxed input) through ultralytics-style
decode / NMS / rescale, the plug-in's int()/shift, sahi's merge and the plug-in's key-point attach.

Comparison rule (north_star: bit-exact slice boxes, kept sets and merge groupings; <= 1e-4 px on coordinates, <= 1e-3 on
scores).  The plug-in truncates float boxes with int() (utils/yolo_wrapper.py:138), which turns Kernel 2's <= 1e-4 px
tolerance into one whole pixel whenever a float coordinate sits within that distance of an integer.  Such "flips" are
identified coordinate by coordinate against the oracle's recorded FLOAT boxes — every other coordinate must be identical —
and no image is skipped: for an image with a flip the merge is checked bit-exactly by running the oracle merge on the
device's own per-slice boxes (identical inputs on both sides); for all others against the pure CPU flow.
"""
from __future__ import annotations

import numpy as np
import torch

BORDERLINE_PX = 2e-4  # K2's coordinate tolerance is 1e-4 px (north_star): inside this distance of an integer, int() may flip


class Recorder:
    def __init__(self):
        self.items = []

    def hook(self, kind, start, x, levels):
        xs = x.float().cpu() if x.dtype != torch.float16 else x.cpu()
        lv = [tuple(t.detach().cpu() for t in l) for l in levels]
        for i in range(x.shape[0]):
            self.items.append((xs[i].contiguous(), [tuple(t[i:i + 1].contiguous() for t in l) for l in lv]))
        return levels

    def lookup(self, im, _levels):
        im0 = im[0] if im.dtype == torch.float16 else im[0].float()
        for x, lv in self.items:
            if x.shape == im0.shape and torch.equal(x, im0):
                return [tuple(t.float() for t in l) for l in lv]
        raise AssertionError("the oracle's letterboxed input has no bit-identical twin among Kernel 1's outputs")


def probe(oracle_yolo_cls):
    """OracleYOLO that also records, per call, its float boxes / scores (before the plug-in's int() truncation)."""

    class Probe(oracle_yolo_cls):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.calls = []
            self.min_gap = 1.0

        def predict(self, *a, **k):
            res = super().predict(*a, **k)
            xy = res[0].boxes.xyxy.double()
            self.calls.append((xy.numpy().copy(), res[0].boxes.conf.numpy().copy()))
            gap = (xy - xy.round()).abs()
            gap = gap[gap > 0]  # exact integers come from clipping to the image bounds, identical on both sides
            if gap.numel():
                self.min_gap = min(self.min_gap, float(gap.min()))
            return res

    return Probe


def as_rows(preds):
    return [([int(v) for v in p.bbox.to_xyxy()], float(p.score.value),
             None if getattr(p, "keypoints", None) is None else np.asarray(p.keypoints, dtype=np.float32)) for p in preds]


def compare_stage1(stage1_rows: np.ndarray, oracle_calls, shifts):
    """Device per-slice detections (rows in the reference's append order) against the oracle's per-call float boxes.
    Returns the number of int() flips; raises on any difference that is not such a flip."""
    want_xy = np.concatenate([c[0].reshape(-1, 4) for c in oracle_calls]) if oracle_calls else np.zeros((0, 4))
    want_sc = np.concatenate([c[1].reshape(-1) for c in oracle_calls]) if oracle_calls else np.zeros((0,))
    want_shift = np.concatenate([np.tile(np.asarray(s, dtype=np.int64), (len(c[0]), 2)) for c, s in zip(oracle_calls, shifts)]) \
        if oracle_calls else np.zeros((0, 4), np.int64)
    assert len(stage1_rows) == len(want_xy), f"per-slice detection count {len(stage1_rows)} vs {len(want_xy)}"
    got = stage1_rows[:, :4].astype(np.int64)
    want = want_xy.astype(np.int64) + want_shift  # astype(int) truncation, then + shift (a10's clamp is a no-op here)
    assert np.abs(stage1_rows[:, 4] - want_sc).max(initial=0.0) <= 1e-3, "per-slice scores differ by more than 1e-3"
    diff = got != want
    if not diff.any():
        return 0
    gap = np.abs(want_xy - np.round(want_xy))
    bad = diff & ~((np.abs(got - want) == 1) & (gap <= BORDERLINE_PX))
    assert not bad.any(), f"per-slice boxes differ beyond an int() flip: got {got[bad.any(1)][:4]} want {want[bad.any(1)][:4]} float {want_xy[bad.any(1)][:4]}"
    return int(diff.sum())


def oracle_merge_of_rows(stage1_rows, ptype, metric, thr):
    """sahi's merge (oracle) applied to the device's own per-slice detections: identical inputs on both sides."""
    from oracle import postprocess as opp
    from oracle.annotation import ObjectPrediction as OracleOP

    preds = [OracleOP(bbox=[int(v) for v in r[:4]], score=float(r[4]), category_id=0, category_name="face") for r in stage1_rows]
    if len(preds) > 1:
        preds = opp.POSTPROCESS_NAME_TO_CLASS[ptype](match_threshold=thr, match_metric=metric, class_agnostic=False)(preds)
    return preds


def compare_keypoints(a, b):
    for ra, rb in zip(a, b):
        assert (ra[2] is None) == (rb[2] is None)
        if ra[2] is not None:
            tol = np.maximum(1e-4, 2 * np.spacing(np.abs(rb[2][:, :2])))
            assert (np.abs(ra[2][:, :2] - rb[2][:, :2]) <= tol).all() and np.abs(ra[2][:, 2] - rb[2][:, 2]).max() <= 1e-3


def to_xywh(rows):
    return np.array([[r[0][0], r[0][1], r[0][2] - r[0][0], r[0][3] - r[0][1], r[1]] for r in rows], dtype=float).reshape(-1, 5)


def run_sliced_case(H, W, sl, ov, imgsz, conf, ptype, metric, n_images, seed0=100, mean_faces=8, face_px=(8, 120),
                    half=True, images=None):
    """Fused device path vs oracle flow over `n_images` synthetic images; returns a summary dict."""
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.sahi_api import get_sliced_prediction
    from fsd_b200.synthetic import make_image
    from fsd_b200.yolo import YOLO
    from oracle import predict as opred
    from oracle import slicing as oslice
    from oracle import widerface_eval as oe
    from oracle.yolo_head import OracleYOLO
    from oracle.yolo_wrapper import YOLOv11PoseDetectionModel as OracleModel
    import fsd_b200.widerface_eval as pe

    det_flag = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True  # reproducible head tensors run to run
    try:
        yolo = YOLO("random-init")
        model = YOLOv11PoseDetectionModel(model=yolo, confidence_threshold=conf, device="cuda:0", image_size=imgsz, half=half)
        eng = model.engine()
        preds_gpu, preds_cpu, preds_replay, gts = [], [], [], []
        n_boxes = n_flips = n_flip_images = n_stage1 = 0
        for i in range(n_images):
            if images is not None:
                img, gt = images[i]
            else:
                img, gt = make_image(seed0 + i, H, W, mean_faces=mean_faces, face_px=face_px)
            rec = Recorder()
            eng.head_hook = rec.hook
            model.keypoints_cache = {}
            got = get_sliced_prediction(img, model, slice_height=sl, slice_width=sl, overlap_height_ratio=ov,
                                        overlap_width_ratio=ov, postprocess_type=ptype, postprocess_match_metric=metric,
                                        postprocess_match_threshold=0.5, verbose=0)
            eng.head_hook = None
            stage1 = eng.last_stage1["rows"][0]
            got_list = model.attach_keypoints_to_predictions(got.object_prediction_list)
            oyolo = probe(OracleYOLO)(None, half=half, head_hook=rec.lookup)
            omodel = OracleModel(model=oyolo, confidence_threshold=conf, device="cpu", image_size=imgsz)
            want = opred.get_sliced_prediction(img, omodel, slice_height=sl, slice_width=sl, overlap_height_ratio=ov,
                                               overlap_width_ratio=ov, postprocess_type=ptype, postprocess_match_metric=metric,
                                               postprocess_match_threshold=0.5, verbose=0)
            want_list = omodel.attach_keypoints_to_predictions(want.object_prediction_list)
            boxes = oslice.get_slice_bboxes(img.shape[0], img.shape[1], sl, sl, True, ov, ov)
            shifts = [b[:2] for b in boxes] + ([[0, 0]] if len(boxes) > 1 else [])
            # (the engine re-runs a batch once when a slice overflowed the candidate capacity: the hook then saw every input twice)
            assert len(oyolo.calls) == len(shifts) and len(rec.items) % len(shifts) == 0 and len(rec.items) > 0, \
                (len(oyolo.calls), len(shifts), len(rec.items))
            flips = compare_stage1(stage1, oyolo.calls, shifts)
            a, b = as_rows(got_list), as_rows(want_list)
            # the merge on identical inputs: always bit-exact
            replay = as_rows(oracle_merge_of_rows(stage1, ptype, metric, 0.5))
            assert [r[0] for r in a] == [r[0] for r in replay], "merged boxes differ from the oracle merge of the same per-slice boxes"
            assert np.allclose([r[1] for r in a], [r[1] for r in replay], atol=0, rtol=0)
            if flips == 0:
                assert list(model.keypoints_cache.keys()) == list(omodel.keypoints_cache.keys())
                assert [r[0] for r in a] == [r[0] for r in b], "merged boxes differ"
                assert np.allclose([r[1] for r in a], [r[1] for r in b], atol=1e-3, rtol=0)
                compare_keypoints(a, b)
            else:
                n_flips += flips
                n_flip_images += 1
            n_boxes += len(a)
            n_stage1 += len(stage1)
            preds_gpu.append(to_xywh(a))
            preds_cpu.append(to_xywh(b))
            preds_replay.append(to_xywh(replay))
            gts.append(gt)
        aps = {}
        for setting in ("easy", "medium", "hard"):
            keeps = [oe.difficulty_keep_lists(g)[setting] for g in gts]
            ap_cpu, _ = oe.evaluate_setting(preds_cpu, gts, keeps, thresh_num=1000)
            ap_replay, _ = oe.evaluate_setting(preds_replay, gts, keeps, thresh_num=1000)
            ap_gpu, _ = pe.evaluate_setting(preds_gpu, gts, keeps, thresh_num=1000)
            assert ap_gpu == ap_replay, f"{setting}: AP {ap_gpu} vs {ap_replay} on identical per-slice boxes"
            if n_flips == 0:
                assert ap_gpu == ap_cpu, f"{setting}: AP {ap_gpu} vs {ap_cpu}"
            else:
                assert abs(ap_gpu - ap_cpu) <= 0.01, f"{setting}: AP {ap_gpu} vs {ap_cpu} with {n_flips} int() flips"
            aps[setting] = (ap_gpu, ap_cpu)
        return dict(boxes=n_boxes, stage1=n_stage1, flips=n_flips, flip_images=n_flip_images, aps=aps)
    finally:
        torch.backends.cudnn.deterministic = det_flag


def merge_gate(device, n=1024, seed=0, merge_type="GREEDYNMM", metric="IOS"):
    """Kernel 3 on one synthetic segment of `n` integer boxes against the sahi oracle: identical keeps and merged boxes."""
    from fsd_b200 import ops
    from oracle import postprocess as opp
    from oracle.annotation import ObjectPrediction as OracleOP

    rng = np.random.default_rng(seed)
    side = int(40 * np.sqrt(n))
    xy = rng.integers(0, side, (n, 2))
    wh = rng.integers(8, 60, (n, 2))
    rows = np.concatenate([xy, xy + wh, rng.uniform(0.3, 1.0, (n, 1)).astype(np.float32), np.zeros((n, 1))], 1).astype(np.float32)
    res = ops.merge_segments(torch.from_numpy(rows).to(device), torch.zeros(1, dtype=torch.int32, device=device), None, n,
                             merge_type=merge_type, metric=metric, thr=0.5, precision="fp64", tie_rule="box_lex")
    k = int(res["keep_count"][0])
    got = res["boxes"][:k].cpu().numpy().astype(np.int64).tolist()
    preds = [OracleOP(bbox=[int(v) for v in r[:4]], score=float(r[4]), category_id=0, category_name="face") for r in rows]
    want = opp.POSTPROCESS_NAME_TO_CLASS[merge_type](match_threshold=0.5, match_metric=metric, class_agnostic=True)(preds)
    assert got == [[int(v) for v in p.bbox.to_xyxy()] for p in want], f"Kernel 3 {merge_type}/{metric} n={n} differs from the oracle"
    return dict(n=n, kept=k)


def esrgan_gate(device, h=97, w=131, scale=2, tile=64):
    """Kernel 4 crop -> exact nearest up-sampler -> stitch against the oracle RealESRGANer: bit-exact."""
    from fsd_b200 import ops
    from oracle.esrgan import NearestUpsampler, RealESRGANer as OracleESRGANer

    img = np.random.default_rng(5).integers(0, 256, (h, w, 3), dtype=np.uint8)
    table, _ = ops.esrgan_tile_table(h, w, scale, tile, 10, 0)
    tiles, tab_dev = ops.esrgan_crop(torch.from_numpy(img).to(device), table, scale, 0, torch.float32)
    outbuf = ops.esrgan_out_buffer(table, scale, torch.float32, device)
    up = NearestUpsampler(scale)
    for row in table:
        ops.tile_view(outbuf, row, scale, out=True).copy_(up(ops.tile_view(tiles, row)))
    got = ops.esrgan_stitch(outbuf, table, tab_dev, scale, h, w).cpu().numpy()
    want, _ = OracleESRGANer(scale=scale, model=up, tile=tile, tile_pad=10, pre_pad=0, half=False).enhance(img, outscale=scale)
    assert np.array_equal(got, want), "Kernel 4 differs from the oracle"
    return dict(tiles=len(table))
