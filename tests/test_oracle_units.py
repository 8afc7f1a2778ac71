"""CPU unit tests of the oracle and of the host-side logic of the C library (no GPU needed)."""
import math

import numpy as np
import pytest
import torch
import torchvision
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import esrgan as oesr
from oracle import letterbox as olb
from oracle import postprocess as opp
from oracle import slicing as osl
from oracle import yolo_head as oy
from oracle.annotation import ObjectPrediction

# ------------------------------------------------------------------------------------------- slicing (App. B)
APP_B = [  # W, H, slice, overlap, n, x starts, y starts
    (1920, 1080, 640, 0.2, 8, [0, 512, 1024, 1280], [0, 440]),
    (1024, 768, 512, 0.2, 6, [0, 410, 512], [0, 256]),
    (3840, 2160, 640, 0.2, 32, [0, 512, 1024, 1536, 2048, 2560, 3072, 3200], [0, 512, 1024, 1520]),
    (3840, 2160, 640, 0.25, 40, [0, 480, 960, 1440, 1920, 2400, 2880, 3200], [0, 480, 960, 1440, 1520]),
    (2048, 1366, 640, 0.25, 12, [0, 480, 960, 1408], [0, 480, 726]),
    (1024, 768, 640, 0.2, 4, [0, 384], [0, 128]),
]


@pytest.mark.parametrize("case", APP_B)
def test_slice_grid_known_answers(case):
    import fsd_b200._cabi as cabi

    W, H, s, ov, n, xs, ys = case
    boxes = osl.get_slice_bboxes(H, W, s, s, overlap_height_ratio=ov, overlap_width_ratio=ov)
    assert len(boxes) == n
    assert sorted({b[0] for b in boxes}) == xs and sorted({b[1] for b in boxes}) == ys
    assert all(b[2] - b[0] == s and b[3] - b[1] == s for b in boxes)
    assert boxes == [[x, y, x + s, y + s] for y in ys for x in xs]  # row-major
    assert cabi.slice_plan(H, W, s, s, ov, ov) == boxes              # the C planner agrees


def test_slice_image_smaller_than_slice():
    import fsd_b200._cabi as cabi

    boxes = osl.get_slice_bboxes(480, 750, 640, 640, overlap_height_ratio=0.2, overlap_width_ratio=0.2)
    assert boxes == [[0, 0, 640, 480], [110, 0, 750, 480]]
    assert cabi.slice_plan(480, 750, 640, 640, 0.2, 0.2) == boxes


@settings(max_examples=150, deadline=None)
@given(st.integers(1, 3000), st.integers(1, 3000), st.integers(16, 1024), st.integers(16, 1024),
       st.sampled_from([0.0, 0.1, 0.2, 0.25, 0.4]), st.sampled_from([0.0, 0.1, 0.2, 0.25, 0.4]))
def test_slice_plan_properties(H, W, sh, sw, oh, ow):
    import fsd_b200._cabi as cabi

    boxes = osl.get_slice_bboxes(H, W, sh, sw, overlap_height_ratio=oh, overlap_width_ratio=ow)
    assert cabi.slice_plan(H, W, sh, sw, oh, ow) == boxes
    cover = np.zeros((H, W), dtype=bool)
    for x0, y0, x1, y1 in boxes:
        assert 0 <= x0 < x1 <= W and 0 <= y0 < y1 <= H
        assert (x1 - x0 == sw or W < sw) and (y1 - y0 == sh or H < sh)
        cover[y0:y1, x0:x1] = True
    assert cover.all()


def test_read_image_as_pil_keeps_ndarray_channel_order():
    arr = np.arange(6 * 7 * 3, dtype=np.uint8).reshape(6, 7, 3)
    assert np.array_equal(np.asarray(osl.read_image_as_pil(arr)), arr)  # no BGR->RGB fix for arrays


# ------------------------------------------------------------------------------------------- letterbox
SHAPES = [((640, 640), 1024), ((512, 512), 1024), ((1080, 1920), 1024), ((768, 1024), 1024), ((1366, 2048), 1024),
          ((480, 640), 1024), ((37, 53), 1024), ((333, 517), 640), ((2160, 3840), 1024), ((600, 800), 512)]


@pytest.mark.parametrize("shape", SHAPES)
def test_letterbox_integer_restatement_equals_cv2(shape):
    import fsd_b200._cabi as cabi

    (h, w), imgsz = shape
    img = np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(olb.letterbox_numpy(img, imgsz), olb.letterbox_cv2(img, imgsz))
    g, c = olb.letterbox_geometry(h, w, imgsz), cabi.letterbox_geometry(h, w, imgsz)
    for k in ("new_w", "new_h", "left", "top", "out_w", "out_h", "mode"):
        assert g[k] == c[k]
    assert g["gain"] == c["gain"]


def test_letterbox_known_answers():
    g = olb.letterbox_geometry(1080, 1920)
    assert (g["out_h"], g["out_w"], g["top"], g["left"]) == (576, 1024, 0, 0) and abs(g["gain"] - 0.53333) < 1e-4
    assert olb.letterbox_geometry(640, 640)["out_h"] == 1024 and olb.letterbox_geometry(768, 1024)["mode"] == 0
    g = olb.letterbox_geometry(480, 640)
    assert (g["out_h"], g["out_w"], g["gain"]) == (768, 1024, 1.6)
    assert olb.letterbox_geometry(1366, 2048)["mode"] == 2  # cv2 takes the exact-2x area path


def test_normalisation_facts_used_by_kernel1():
    v = torch.arange(256, dtype=torch.uint8)
    assert torch.equal(v.half() / 255, (v.float() / 255).half())                  # torch's half division
    assert torch.equal((v.float() * np.float32(1 / 255)).half(), v.half() / 255)  # one multiply is enough for fp16
    q = v.float() * np.float32(1.0 / 255.0)                                       # fp32 needs the Newton correction
    r = (v.double() - q.double() * 255.0).float()
    fixed = (r.double() * float(np.float32(1.0 / 255.0)) + q.double()).float()
    assert torch.equal(fixed, v.float() / 255)


# ------------------------------------------------------------------------------------------- head decode / NMS
def test_make_anchors_and_decode_shapes():
    a, s = oy.make_anchors([(4, 5), (2, 3), (1, 2)])
    assert a.shape == (2, 28) and s.shape == (1, 28)
    assert a[:, 0].tolist() == [0.5, 0.5] and a[:, 6].tolist() == [1.5, 1.5] and s[0, 20].item() == 16
    g = torch.Generator().manual_seed(0)
    levels = [(torch.randn(2, 64, h, w, generator=g), torch.randn(2, 1, h, w, generator=g), torch.randn(2, 15, h, w, generator=g))
              for h, w in [(8, 8), (4, 4), (2, 2)]]
    y = oy.decode_head(levels)
    assert y.shape == (2, 20, 84) and (y[:, 4] >= 0).all() and (y[:, 4] <= 1).all()
    # a one-hot DFL (bin 3 on every side) gives a box of exactly 6 strides around the anchor
    box = torch.full((1, 64, 1, 1), -100.0)
    box[0, [3, 19, 35, 51]] = 100.0
    y = oy.decode_head([(box, torch.zeros(1, 1, 1, 1), torch.zeros(1, 15, 1, 1))] * 1 + [(box, torch.zeros(1, 1, 1, 1), torch.zeros(1, 15, 1, 1))] * 2)
    assert y[0, :4, 0].tolist() == [4.0, 4.0, 48.0, 48.0] and y[0, 4, 0].item() == 0.5


def test_nms_restatement_uses_torchvision_rule():
    pred = torch.zeros((1, 20, 3))
    pred[0, :4, 0] = torch.tensor([50.0, 50, 40, 40])
    pred[0, :4, 1] = torch.tensor([52.0, 50, 40, 40])   # IoU ~0.9 with box 0
    pred[0, :4, 2] = torch.tensor([200.0, 200, 40, 40])
    pred[0, 4] = torch.tensor([0.9, 0.8, 0.3])
    out = oy.non_max_suppression(pred, conf_thres=0.25)[0]
    assert out.shape[0] == 2 and out[:, 4].tolist() == pytest.approx([0.9, 0.3])
    assert oy.non_max_suppression(pred, conf_thres=0.95)[0].shape[0] == 0


def test_scale_boxes_roundtrip_1080p():
    g = olb.letterbox_geometry(1080, 1920)
    b = torch.tensor([[100.0, 50.0, 300.0, 250.0]])
    out = oy.scale_boxes((g["out_h"], g["out_w"]), b.clone(), (1080, 1920))
    assert torch.allclose(out, b / np.float32(g["gain"]))


# ------------------------------------------------------------------------------------------- merge (hand-made)
def P(box, score, cat=0):
    return ObjectPrediction(bbox=box, score=score, category_id=cat, category_name=str(cat))


def test_threshold_equality_is_dropped_not_merged():
    a, b = P([0, 0, 100, 100], 0.9), P([0, 0, 100, 50], 0.8)  # IOU exactly 0.5
    out = opp.GreedyNMMPostprocess(0.5, "IOU", True)([a, b])
    assert len(out) == 1 and out[0].bbox.to_xyxy() == [0, 0, 100, 100]
    assert [o.bbox.to_xyxy() for o in opp.NMSPostprocess(0.5, "IOU", True)([a, b])] == [[0, 0, 100, 100]]
    out = opp.GreedyNMMPostprocess(0.49, "IOU", True)([a, b])
    assert out[0].bbox.to_xyxy() == [0, 0, 100, 100] and out[0].score.value == 0.9


def test_zero_area_and_equal_scores():
    preds = [P([10, 10, 10, 40], 0.7), P([0, 0, 50, 50], 0.7), P([5, 5, 45, 45], 0.7)]
    pp = opp.GreedyNMMPostprocess(0.5, "IOS", True, tie_rule="index")
    out = pp(preds)
    assert pp.last_keep_to_merge == {0: [], 1: [2]} and [o.bbox.to_xyxy() for o in out] == [[10, 10, 10, 40], [0, 0, 50, 50]]
    # variant N (lines 606-617), the default: box 1 must not test box 2 (equal score, lexicographically larger
    # tuple), so box 2 is kept as well — and, visited later, claims the already kept box 1 into its own merge list
    pp = opp.GreedyNMMPostprocess(0.5, "IOS", True)
    out = pp(preds)
    assert pp.last_keep_to_merge == {0: [], 1: [], 2: [1]}
    assert [o.bbox.to_xyxy() for o in out] == [[10, 10, 10, 40], [0, 0, 50, 50], [0, 0, 50, 50]]
    assert opp.nms(opp.to_array(preds), "IOS", 0.5) == [0, 1, 2] and opp.nms(opp.to_array(preds), "IOS", 0.5, tie_rule="index") == [0, 1]
    # equal scores, lexicographically SMALLER later box: the plain rule applies
    rev = [P([5, 5, 45, 45], 0.7), P([0, 0, 50, 50], 0.7)]
    assert opp.greedy_nmm(opp.to_array(rev), "IOS", 0.5) == {0: [1]}
    # a chain inside one tie group: 0 <- 1 <- 2 are each claimed by the next (lexicographically larger) keep
    chain = [P([0, 0, 40, 40], 0.5), P([1, 0, 41, 40], 0.5), P([2, 0, 42, 40], 0.5)]
    pp = opp.GreedyNMMPostprocess(0.5, "IOU", True)
    out = pp(chain)
    assert pp.last_keep_to_merge == {0: [], 1: [0], 2: [1]}
    assert [o.bbox.to_xyxy() for o in out] == [[0, 0, 40, 40], [0, 0, 41, 40], [0, 0, 42, 40]]  # 2 merges the MERGED box 1


def test_chain_greedy_vs_transitive():
    preds = [P([0, 0, 100, 100], 0.9), P([60, 0, 160, 100], 0.8), P([120, 0, 220, 100], 0.7)]
    g = opp.GreedyNMMPostprocess(0.2, "IOU", True)
    out = g(preds)
    assert g.last_keep_to_merge == {0: [1], 2: []} and [o.bbox.to_xyxy() for o in out] == [[0, 0, 160, 100], [120, 0, 220, 100]]
    n = opp.NMMPostprocess(0.2, "IOU", True)
    out = n(preds)
    assert n.last_keep_to_merge == {0: [1, 2]}
    # has_match(A u B, C) is evaluated against the GROWN box: IoU([0,160] , [120,220]) = 40/220 < 0.2 -> C is dropped
    assert [o.bbox.to_xyxy() for o in out] == [[0, 0, 160, 100]]


def test_union_growth_changes_later_match():
    preds = [P([0, 0, 40, 40], 0.9), P([0, 0, 80, 40], 0.8), P([50, 0, 80, 40], 0.7)]
    out = opp.GreedyNMMPostprocess(0.5, "IOS", True)(preds)
    # C does not touch A, is claimed by nothing in the tensor pass ... B is merged into A, C stays its own keep
    assert [o.bbox.to_xyxy() for o in out] == [[0, 0, 80, 40], [50, 0, 80, 40]]


def test_batched_variants_respect_categories():
    preds = [P([0, 0, 50, 50], 0.9, 0), P([1, 1, 50, 50], 0.8, 1), P([2, 2, 50, 50], 0.7, 0)]
    assert len(opp.NMSPostprocess(0.5, "IOU", False)(preds)) == 2
    assert len(opp.NMSPostprocess(0.5, "IOU", True)(preds)) == 1
    with pytest.raises(NotImplementedError):
        opp.LSNMSPostprocess()(preds)


# ------------------------------------------------------------------------------------------- ESRGAN tiling
def test_esrgan_tile_table_known_answers():
    import fsd_b200.ops as ops

    up = oesr.RealESRGANer(scale=2, model=oesr.NearestUpsampler(2), tile=400, tile_pad=10, pre_pad=0)
    up.pre_process(np.zeros((1080, 1920, 3), np.float32))
    t = up.tile_table()
    assert len(t) == 15 and t[0][:4] == (0, 0, 410, 410) and t[1][:4] == (390, 0, 810, 410)
    assert t[4][:4] == (1590, 0, 1920, 410) and t[14][:4] == (1590, 790, 1920, 1080)
    tab, hw = ops.esrgan_tile_table(1080, 1920, 2, 400, 10, 0)
    assert hw == (1080, 1920)
    assert [tuple(int(v) for v in (r[0], r[1], r[0] + r[2], r[1] + r[3], r[4], r[5], r[4] + r[6], r[5] + r[7])) for r in tab] == t
    assert ops.esrgan_tile_table(1079, 1919, 2, 400)[1] == (1080, 1920)  # mod-pad to even for x2
    assert ops.esrgan_tile_table(1079, 1919, 4, 400)[1] == (1079, 1919)


@pytest.mark.parametrize("size", [(1079, 1919), (37, 53), (600, 500)])
@pytest.mark.parametrize("tile", [256, 400, 200])
@pytest.mark.parametrize("scale", [2, 4])
def test_esrgan_identity_upsampler_roundtrip(size, tile, scale):
    img = np.random.default_rng(size[0]).integers(0, 256, (*size, 3), dtype=np.uint8)
    out, _ = oesr.RealESRGANer(scale=scale, model=oesr.NearestUpsampler(scale), tile=tile, tile_pad=10, pre_pad=0).enhance(img, outscale=scale)
    assert np.array_equal(out, np.repeat(np.repeat(img, scale, 0), scale, 1))
