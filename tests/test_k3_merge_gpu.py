"""Kernel 3 parity: batched NMS / GREEDYNMM / NMM vs torchvision.ops.nms (stage-1 rule) and the sahi restatement
(oracle/postprocess.py, stage-2 rule).  Kept-index lists, merge groupings and merged boxes must be IDENTICAL."""
import numpy as np
import pytest
import torch
import torchvision

from oracle import postprocess as opp
from oracle.annotation import ObjectPrediction

pytestmark = pytest.mark.gpu


def sahi_like_boxes(rng, n_faces, dup=(1, 5), size=(8, 200), canvas=(1920, 1080), ncat=1, score_ties=False):
    """int boxes with jittered / truncated duplicates, like overlapping SAHI slices produce."""
    rows = []
    for _ in range(n_faces):
        w = int(np.exp(rng.uniform(np.log(size[0]), np.log(size[1]))))
        h = int(w * rng.uniform(0.8, 1.4))
        x = int(rng.integers(0, canvas[0] - w))
        y = int(rng.integers(0, canvas[1] - h))
        cat = int(rng.integers(0, ncat))
        for _ in range(int(rng.integers(dup[0], dup[1] + 1))):
            j = rng.integers(-3, 4, size=4)
            x1, y1, x2, y2 = x + j[0], y + j[1], x + w + j[2], y + h + j[3]
            if rng.random() < 0.3:  # slice-truncated copy
                if rng.random() < 0.5:
                    x2 = x1 + max(1, int((x2 - x1) * rng.uniform(0.3, 0.9)))
                else:
                    y1 = y2 - max(1, int((y2 - y1) * rng.uniform(0.3, 0.9)))
            s = rng.uniform(0.3, 1.0)
            if score_ties:
                s = round(s, 1)
            rows.append([max(x1, 0), max(y1, 0), max(x2, 1), max(y2, 1), np.float32(s), cat])
    return np.array(rows, dtype=np.float32).reshape(-1, 6)


def run_kernel(dev, segs, **kw):
    """segs: list of [n,6] arrays -> per-segment dicts of numpy results (local indices)."""
    import fsd_b200.ops as ops

    counts = [len(s) for s in segs]
    cap = max(max(counts), 1)
    offs = np.arange(len(segs)) * cap
    rows = np.zeros((len(segs) * cap, 6), dtype=np.float32)
    for i, s in enumerate(segs):
        rows[offs[i]: offs[i] + len(s)] = s
    t = torch.from_numpy(rows).to(dev)
    cats = t[:, 5].to(torch.int32).contiguous()
    res = ops.merge_segments(t, torch.tensor(offs, dtype=torch.int32, device=dev),
                             torch.tensor(counts, dtype=torch.int32, device=dev), cap, cats=cats, **kw)
    out = []
    kc = res["keep_count"].cpu().numpy()
    for i, s in enumerate(segs):
        o, k = int(offs[i]), int(kc[i])
        parent = res["parent"].cpu().numpy()[o: o + len(s)]
        out.append(dict(keep=(res["keep"].cpu().numpy()[o: o + k] - o).tolist(),
                        parent=np.where(parent >= 0, parent - o, -1),
                        boxes=res["boxes"].cpu().numpy()[o: o + k], scores=res["scores"].cpu().numpy()[o: o + k],
                        cats=res["cats"].cpu().numpy()[o: o + k]))
    return out


def groups_from_parent(parent, keep):
    g = {k: [] for k in keep}
    for j, p in enumerate(parent):
        if p >= 0 and p != j:
            g[int(p)].append(j)
    return g


# ---------------------------------------------------------------------------------------------- stage 1
@pytest.mark.parametrize("sizes", [[0, 1, 2, 63, 64, 65, 300], [1000, 17], [5000]])
def test_stage1_matches_torchvision_nms(cuda_device, sizes):
    rng = np.random.default_rng(5)
    segs = []
    for n in sizes:
        xy = rng.uniform(0, 900, (n, 2)).astype(np.float32)
        wh = rng.uniform(4, 160, (n, 2)).astype(np.float32)
        sc = rng.uniform(0.01, 1, n).astype(np.float16).astype(np.float32)  # fp16-rounded => many exact ties
        segs.append(np.concatenate([xy, xy + wh, sc[:, None], np.zeros((n, 1), np.float32)], 1))
    got = run_kernel(cuda_device, segs, merge_type="NMS", metric="IOU", thr=0.7, cmp_strict=True, precision="fp32",
                     class_agnostic=True, pre_cap=30000, max_keep=300)
    for s, g in zip(segs, got):
        t = torch.from_numpy(s)
        ref = torchvision.ops.nms(t[:, :4], t[:, 4], 0.7)[:300].tolist() if len(s) else []
        assert g["keep"] == ref
        if len(s):
            assert np.array_equal(g["boxes"], s[ref, :4]) and np.array_equal(g["scores"], s[ref, 4])


def test_stage1_pre_cap(cuda_device):
    rng = np.random.default_rng(9)
    n = 700
    xy = rng.uniform(0, 3000, (n, 2)).astype(np.float32)
    s = np.concatenate([xy, xy + 5, rng.uniform(0, 1, (n, 1)).astype(np.float32), np.zeros((n, 1), np.float32)], 1)
    got = run_kernel(cuda_device, [s], merge_type="NMS", metric="IOU", thr=0.7, cmp_strict=True, precision="fp32",
                     pre_cap=500, max_keep=0)[0]
    t = torch.from_numpy(s)
    top = t[:, 4].argsort(descending=True, stable=True)[:500]
    ref = top[torchvision.ops.nms(t[top, :4], t[top, 4], 0.7)].tolist()
    assert got["keep"] == ref
    assert (got["parent"] == -1).sum() == 200


# ---------------------------------------------------------------------------------------------- stage 2
def _oracle(seg, ptype, metric, thr, agnostic, tie_rule="box_lex"):
    preds = [ObjectPrediction(bbox=[int(v) for v in r[:4]], score=float(r[4]), category_id=int(r[5]),
                              category_name=str(int(r[5]))) for r in seg]
    pp = opp.POSTPROCESS_NAME_TO_CLASS[ptype](match_threshold=thr, match_metric=metric, class_agnostic=agnostic, tie_rule=tie_rule)
    out = pp(preds) if len(preds) else []
    if ptype == "NMS":
        keep, groups = (pp.last_keep if len(preds) else []), None
    else:
        ktm = pp.last_keep_to_merge if len(preds) else {}
        keep, groups = list(ktm.keys()), ktm
    return keep, groups, out


@pytest.mark.parametrize("ptype", ["NMS", "GREEDYNMM", "NMM"])
@pytest.mark.parametrize("metric", ["IOS", "IOU"])
@pytest.mark.parametrize("thr", [0.3, 0.5, 0.7])
@pytest.mark.parametrize("tie_rule", ["box_lex", "index"])
def test_stage2_matches_sahi_oracle(cuda_device, ptype, metric, thr, tie_rule):
    rng = np.random.default_rng({"NMS": 1, "GREEDYNMM": 2, "NMM": 3}[ptype] * 100 + (7 if metric == "IOS" else 0) + int(thr * 10))
    segs = [sahi_like_boxes(rng, nf) for nf in (0, 1, 3, 40, 150)]
    segs.append(sahi_like_boxes(rng, 60, score_ties=True))       # scores rounded to one decimal: large tie groups
    segs.append(sahi_like_boxes(rng, 400, dup=(2, 6), size=(20, 120), canvas=(1200, 800), score_ties=True))  # dense + ties, > 1024 boxes
    got = run_kernel(cuda_device, segs, merge_type=ptype, metric=metric, thr=thr, cmp_strict=False, precision="fp64",
                     class_agnostic=True, tie_rule=tie_rule)
    for seg, g in zip(segs, got):
        keep, groups, out = _oracle(seg, ptype, metric, thr, True, tie_rule)
        assert g["keep"] == keep
        if groups is not None:
            gg = groups_from_parent(g["parent"], g["keep"])
            assert {k: sorted(v) for k, v in gg.items()} == {k: sorted(v) for k, v in groups.items()}
        ref_boxes = np.array([o.bbox.to_xyxy() for o in out], dtype=np.float32).reshape(-1, 4)
        assert np.array_equal(g["boxes"], ref_boxes)
        assert np.array_equal(g["scores"], np.array([o.score.value for o in out], dtype=np.float32))


@pytest.mark.parametrize("ptype", ["NMS", "GREEDYNMM", "NMM"])
def test_stage2_per_category(cuda_device, ptype):
    rng = np.random.default_rng(77)
    seg = sahi_like_boxes(rng, 80, ncat=3)
    g = run_kernel(cuda_device, [seg], merge_type=ptype, metric="IOS", thr=0.5, cmp_strict=False, precision="fp64",
                   class_agnostic=False, tie_rule="box_lex")[0]
    keep, groups, out = _oracle(seg, ptype, "IOS", 0.5, False)
    # the kernel emits one score-descending list; sahi's batched_* variants list category by category
    assert sorted(g["keep"]) == sorted(keep)
    by_keep = {k: (b, c) for k, b, c in zip(g["keep"], g["boxes"], g["cats"])}
    for k, o in zip(keep, out):
        assert np.array_equal(by_keep[k][0], np.array(o.bbox.to_xyxy(), dtype=np.float32))
        assert by_keep[k][1] == o.category.id
    if groups is not None:
        gg = groups_from_parent(g["parent"], g["keep"])
        assert {k: sorted(v) for k, v in gg.items()} == {k: sorted(v) for k, v in groups.items()}


def test_hand_made_edge_cases(cuda_device):
    # metric == threshold exactly: in the merge list (>=) but NOT merged (has_match is strict) -> it disappears
    a = [0, 0, 100, 100, 0.9, 0]
    b = [0, 0, 100, 50, 0.8, 0]  # IOU = 0.5 exactly, IOS = 1.0
    seg = np.array([a, b], dtype=np.float32)
    g = run_kernel(cuda_device, [seg], merge_type="GREEDYNMM", metric="IOU", thr=0.5, precision="fp64")[0]
    assert g["keep"] == [0] and g["parent"].tolist() == [0, 0] and g["boxes"].tolist() == [[0, 0, 100, 100]]
    g = run_kernel(cuda_device, [seg], merge_type="NMS", metric="IOU", thr=0.5, precision="fp64")[0]
    assert g["keep"] == [0]
    g = run_kernel(cuda_device, [seg], merge_type="NMS", metric="IOU", thr=0.5, cmp_strict=True, precision="fp32")[0]
    assert g["keep"] == [0, 1]  # torchvision rule: strictly greater
    # zero-area box never matches; equal scores keep input order (plain rule)
    seg = np.array([[10, 10, 10, 40, 0.7, 0], [0, 0, 50, 50, 0.7, 0], [5, 5, 45, 45, 0.7, 0]], dtype=np.float32)
    g = run_kernel(cuda_device, [seg], merge_type="GREEDYNMM", metric="IOS", thr=0.5, precision="fp64")[0]
    assert g["keep"] == [0, 1] and g["parent"].tolist() == [0, 1, 1]
    # sahi 0.11.34's equal-score rule (SURVEY A.2.4 variant N): box 1 does not test the lexicographically larger box 2, so
    # both are kept; box 2, visited later, claims the already kept box 1 and merges it (its output equals box 1's)
    g = run_kernel(cuda_device, [seg], merge_type="GREEDYNMM", metric="IOS", thr=0.5, precision="fp64", tie_rule="box_lex")[0]
    assert g["keep"] == [0, 1, 2] and g["parent"].tolist() == [0, 2, 2]
    assert g["boxes"].tolist() == [[10, 10, 10, 40], [0, 0, 50, 50], [0, 0, 50, 50]]
    g = run_kernel(cuda_device, [seg], merge_type="NMS", metric="IOS", thr=0.5, precision="fp64", tie_rule="box_lex")[0]
    assert g["keep"] == [0, 1, 2]
    # a chain inside one tie group: each keep is claimed by the next one, which folds the MERGED box of its predecessor
    seg = np.array([[0, 0, 40, 40, 0.5, 0], [1, 0, 41, 40, 0.5, 0], [2, 0, 42, 40, 0.5, 0]], dtype=np.float32)
    keep, groups, out = _oracle(seg, "GREEDYNMM", "IOU", 0.5, True)
    g = run_kernel(cuda_device, [seg], merge_type="GREEDYNMM", metric="IOU", thr=0.5, precision="fp64", tie_rule="box_lex")[0]
    assert g["keep"] == keep == [0, 1, 2] and g["boxes"].tolist() == [o.bbox.to_xyxy() for o in out] == [[0, 0, 40, 40], [0, 0, 41, 40], [0, 0, 42, 40]]
    # chain A~B~C, A!~C: greedy keeps A(+B) and C; NMM pulls C in through B
    seg = np.array([[0, 0, 100, 100, 0.9, 0], [60, 0, 160, 100, 0.8, 0], [120, 0, 220, 100, 0.7, 0]], dtype=np.float32)
    keep_g, groups_g, out_g = _oracle(seg, "GREEDYNMM", "IOU", 0.2, True)
    keep_n, groups_n, out_n = _oracle(seg, "NMM", "IOU", 0.2, True)
    gg = run_kernel(cuda_device, [seg], merge_type="GREEDYNMM", metric="IOU", thr=0.2, precision="fp64")[0]
    gn = run_kernel(cuda_device, [seg], merge_type="NMM", metric="IOU", thr=0.2, precision="fp64")[0]
    assert gg["keep"] == keep_g == [0, 2] and gn["keep"] == keep_n == [0]
    assert gg["boxes"].tolist() == [o.bbox.to_xyxy() for o in out_g]
    assert gn["boxes"].tolist() == [o.bbox.to_xyxy() for o in out_n]
    # union growth flips a later has_match: fold order must be rank order
    seg = np.array([[0, 0, 40, 40, 0.9, 0], [0, 0, 80, 40, 0.8, 0], [50, 0, 80, 40, 0.7, 0]], dtype=np.float32)
    for ptype in ("GREEDYNMM", "NMM"):
        keep, groups, out = _oracle(seg, ptype, "IOS", 0.5, True)
        g = run_kernel(cuda_device, [seg], merge_type=ptype, metric="IOS", thr=0.5, precision="fp64")[0]
        assert g["keep"] == keep and g["boxes"].tolist() == [o.bbox.to_xyxy() for o in out]


@pytest.mark.parametrize("n_faces", [400, 2500])
def test_dense_stress_nms_iou(cuda_device, n_faces):
    """config 3: > 1000 boxes per image, NMS / IOU / 0.5 (utils/insightface_wrapper.py contract: int boxes)."""
    rng = np.random.default_rng(31)
    seg = sahi_like_boxes(rng, n_faces, dup=(1, 4), size=(10, 40), canvas=(3840, 2160))
    assert len(seg) > 1000 or n_faces < 1000
    for ptype in ("NMS", "GREEDYNMM"):
        g = run_kernel(cuda_device, [seg], merge_type=ptype, metric="IOU", thr=0.5, precision="fp64", tie_rule="box_lex")[0]
        keep, groups, out = _oracle(seg, ptype, "IOU", 0.5, True)
        assert g["keep"] == keep
        assert np.array_equal(g["boxes"], np.array([o.bbox.to_xyxy() for o in out], dtype=np.float32))


def test_capacity_limit_is_an_error(cuda_device):
    import fsd_b200._cabi as cabi
    import fsd_b200.ops as ops

    t = torch.zeros((8, 6), device=cuda_device)
    with pytest.raises(cabi.FsdError):
        ops.merge_segments(t, torch.zeros(1, dtype=torch.int32, device=cuda_device), None, 40000)


@pytest.mark.parametrize("ptype,metric,prec", [("NMS", "IOU", "fp64"), ("GREEDYNMM", "IOS", "fp64"), ("GREEDYNMM", "IOU", "fp64"),
                                               ("NMS", "IOU", "fp32")])
def test_cluster_path_equals_single_cta_path(cuda_device, ptype, metric, prec, monkeypatch):
    """Segments above 4096 boxes run on a cluster of 8 CTAs (k3_merge_cluster_kernel); keeps, merge parents, merged boxes,
    scores and categories must be identical to the single-CTA kernel on the same input — ragged segment sizes in one
    launch (empty, tiny, just above the shared-memory limit, config 3's 9900), two categories, class-aware."""
    rng = np.random.default_rng(77)
    segs = []
    for n in (0, 5, 4097, 9900, 700):
        seg = sahi_like_boxes(rng, max(3, n // 2), dup=(1, 4), size=(10, 60), canvas=(2560, 1440))[:n] if n else np.zeros((0, 6), np.float32)
        if len(seg):
            seg[:, 5] = rng.integers(0, 2, len(seg))
        segs.append(seg.astype(np.float32))
    kw = dict(merge_type=ptype, metric=metric, thr=0.5, precision=prec, class_agnostic=False)
    if prec == "fp32":
        kw.update(cmp_strict=True, max_keep=300, pre_cap=30000)
    multi = run_kernel(cuda_device, segs, **kw)
    monkeypatch.setenv("FSD_K3_SINGLE_CTA", "1")
    single = run_kernel(cuda_device, segs, **kw)
    for a, b, seg in zip(multi, single, segs):
        assert a["keep"] == b["keep"]
        assert np.array_equal(a["parent"], b["parent"])
        assert np.array_equal(a["boxes"], b["boxes"]) and np.array_equal(a["scores"], b["scores"]) and np.array_equal(a["cats"], b["cats"])
    assert len(multi[3]["keep"]) > 300 or prec == "fp32"


@pytest.mark.parametrize("ptype", ["NMS", "GREEDYNMM"])
def test_cluster_path_with_score_ties_matches_oracle(cuda_device, ptype):
    """> 4096 boxes with scores rounded to two decimals (thousands of exact ties between overlapping boxes): the cluster kernel's
    equal-score rule, backward claims and deferred folds against the ORACLE."""
    rng = np.random.default_rng(123)
    seg = sahi_like_boxes(rng, 2600, dup=(1, 4), size=(10, 60), canvas=(2560, 1440))[:6000]
    seg[:, 4] = np.round(seg[:, 4], 2)
    assert len(seg) > 4096
    g = run_kernel(cuda_device, [seg], merge_type=ptype, metric="IOS", thr=0.5, precision="fp64", tie_rule="box_lex")[0]
    keep, groups, out = _oracle(seg, ptype, "IOS", 0.5, True)
    assert g["keep"] == keep
    assert np.array_equal(g["boxes"], np.array([o.bbox.to_xyxy() for o in out], dtype=np.float32))
    if groups is not None:
        gg = groups_from_parent(g["parent"], g["keep"])
        assert {k: sorted(v) for k, v in gg.items()} == {k: sorted(v) for k, v in groups.items()}
        assert sum(1 for k in keep for j in groups[k] if j in groups) > 0, "the case must contain a keep claimed by a later keep"


def test_null_parent_with_pre_cap(cuda_device):
    """ranks cut by pre_cap must not be written through a NULL parent pointer (engine._stage1 passes none)."""
    import fsd_b200.ops as ops

    rng = np.random.default_rng(4)
    n = 900
    xy = rng.uniform(0, 3000, (n, 2)).astype(np.float32)
    rows = torch.from_numpy(np.concatenate([xy, xy + 9, rng.uniform(0, 1, (n, 1)).astype(np.float32), np.zeros((n, 1), np.float32)], 1)).to(cuda_device)
    res = ops.merge_segments(rows, torch.zeros(1, dtype=torch.int32, device=cuda_device), None, n, merge_type="NMS", metric="IOU",
                             thr=0.7, cmp_strict=True, precision="fp32", pre_cap=500, max_keep=300, want_parent=False)
    torch.cuda.synchronize()
    assert int(res["keep_count"][0]) == 300 and res["parent"] is None


def test_small_segments_in_a_large_capacity_launch(cuda_device):
    """The launch is sized by the capacity (9900 -> cluster-capable), the path by each segment's actual count: tiny, medium and
    > 4096-box segments in ONE call (what engine.detect does for configs 3 / 5 with few detections) all match the oracle."""
    rng = np.random.default_rng(8)
    segs = [sahi_like_boxes(rng, nf, dup=(1, 4), size=(10, 60), canvas=(2560, 1440)) for nf in (2, 30, 700)]
    segs.append(sahi_like_boxes(rng, 2500, dup=(1, 4), size=(10, 60), canvas=(2560, 1440))[:5000])
    import fsd_b200.ops as ops

    cap = 9900
    rows = np.zeros((len(segs) * cap, 6), dtype=np.float32)
    for i, sg in enumerate(segs):
        rows[i * cap: i * cap + len(sg)] = sg
    res = ops.merge_segments(torch.from_numpy(rows).to(cuda_device), torch.arange(len(segs), dtype=torch.int32, device=cuda_device) * cap,
                             torch.tensor([len(sg) for sg in segs], dtype=torch.int32, device=cuda_device), cap, merge_type="GREEDYNMM",
                             metric="IOS", thr=0.5, precision="fp64", tie_rule="box_lex")
    kc = res["keep_count"].cpu().numpy()
    for i, sg in enumerate(segs):
        keep, groups, out = _oracle(sg, "GREEDYNMM", "IOS", 0.5, True)
        assert (res["keep"].cpu().numpy()[i * cap: i * cap + kc[i]] - i * cap).tolist() == keep
        assert np.array_equal(res["boxes"].cpu().numpy()[i * cap: i * cap + kc[i]], np.array([o.bbox.to_xyxy() for o in out], dtype=np.float32))
