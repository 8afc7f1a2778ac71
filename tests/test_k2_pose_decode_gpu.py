"""Kernel 2 parity: fused pose-head decode / finalize vs the torch-CPU restatement of ultralytics (oracle/yolo_head.py).

Tolerances (BASELINE.json north_star): <= 1e-4 px on box / key-point coordinates, <= 1e-3 on scores; the survivor
SET must be identical except for anchors whose oracle score lies within 1e-6 of the threshold (different exp()
implementations may round such a score to either side)."""
import numpy as np
import pytest
import torch

from oracle import yolo_head as oy

pytestmark = pytest.mark.gpu

PX_TOL, SCORE_TOL, BORDER = 1e-4, 1e-3, 1e-6


def px_close(got, ref):
    """<= 1e-4 px, or 2 fp32 ulps of the coordinate where fp32 cannot resolve 1e-4 px (|x| >= 512)."""
    got, ref = np.asarray(got, dtype=np.float32), np.asarray(ref, dtype=np.float32)
    tol = np.maximum(PX_TOL, 2 * np.spacing(np.abs(ref)))
    return bool((np.abs(got - ref) <= tol).all())


def _random_levels(B, H, W, dtype, seed, channels_last=False, cls_mean=-4.0):
    g = torch.Generator().manual_seed(seed)
    levels = []
    for s in (8, 16, 32):
        h, w = H // s, W // s
        box = torch.randn((B, 64, h, w), generator=g) * 1.5 + 1.0
        cls = torch.randn((B, 1, h, w), generator=g) * 2.0 + cls_mean
        kpt = torch.randn((B, 15, h, w), generator=g)
        lv = [t.to(dtype) for t in (box, cls, kpt)]
        levels.append(tuple(lv))
    return levels


def _to_dev(levels, dev, channels_last):
    out = []
    for lv in levels:
        ts = [t.to(dev) for t in lv]
        if channels_last:
            ts = [t.contiguous(memory_format=torch.channels_last) for t in ts]
        out.append(tuple(ts))
    return out


def _oracle_candidates(levels, conf):
    y = oy.decode_head(levels)  # [B,20,A], xywh
    y = y.transpose(1, 2).clone()
    y[..., :4] = oy.xywh2xyxy(y[..., :4])
    return y  # [B,A,20] xyxy, score, 15 kpts


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("conf", [0.5, 0.01])
def test_decode_matches_oracle(cuda_device, dtype, channels_last, conf):
    import fsd_b200.ops as ops

    B, H, W = 3, 256, 320
    levels = _random_levels(B, H, W, dtype, seed=11)
    ref = _oracle_candidates(levels, conf)
    cand, count = ops.pose_decode(_to_dev(levels, cuda_device, channels_last), conf, cap_per_entry=4096)
    cand, count = cand.cpu(), count.cpu()
    n_checked = 0
    for b in range(B):
        n = int(count[b])
        rows = cand[b, :n]
        anchors = rows[:, 5].view(torch.int32).long()
        assert len(set(anchors.tolist())) == n, "duplicate anchors in the candidate list"
        sc = ref[b, :, 4]
        must = set(torch.nonzero(sc > conf + BORDER).flatten().tolist())
        may = set(torch.nonzero(sc > conf - BORDER).flatten().tolist())
        got = set(anchors.tolist())
        assert must <= got <= may, f"survivor set differs: missing {sorted(must - got)[:5]} extra {sorted(got - may)[:5]}"
        r = ref[b, anchors]
        assert px_close(rows[:, :4], r[:, :4])
        assert (rows[:, 4] - r[:, 4]).abs().max() <= SCORE_TOL
        k_got, k_ref = rows[:, 6:21].view(-1, 5, 3), r[:, 5:20].view(-1, 5, 3)
        assert px_close(k_got[..., :2], k_ref[..., :2])
        assert (k_got[..., 2] - k_ref[..., 2]).abs().max() <= SCORE_TOL
        n_checked += n
    assert n_checked > 20, "test data produced too few survivors to mean anything"


def test_capacity_overflow_is_reported(cuda_device):
    import fsd_b200.ops as ops

    levels = _random_levels(1, 128, 128, torch.float32, seed=3, cls_mean=2.0)  # most anchors pass
    cand, count = ops.pose_decode(_to_dev(levels, cuda_device, False), 0.5, cap_per_entry=16)
    assert int(count[0]) > 16  # caller sees that candidates were dropped
    assert torch.isfinite(cand[0, :16, :5]).all()


def test_empty_result(cuda_device):
    import fsd_b200.ops as ops

    levels = _random_levels(2, 64, 64, torch.float16, seed=5, cls_mean=-30.0)
    _, count = ops.pose_decode(_to_dev(levels, cuda_device, False), 0.5, cap_per_entry=8)
    assert count.tolist() == [0, 0]


@pytest.mark.parametrize("shape", [((640, 640), 1024), ((512, 512), 1024), ((1080, 1920), 1024), ((480, 750), 1024)])
def test_stage1_nms_and_finalize_match_oracle(cuda_device, shape):
    """decode -> per-entry NMS (Kernel 3, torchvision rule) -> finalize, against OracleYOLO + the plugin's int()/shift."""
    import fsd_b200._cabi as cabi
    import fsd_b200.ops as ops
    from oracle import letterbox as olb

    (src_h, src_w), imgsz = shape
    g = olb.letterbox_geometry(src_h, src_w, imgsz)
    B, conf = 2, 0.25
    levels = _random_levels(B, g["out_h"], g["out_w"], torch.float32, seed=21, cls_mean=-5.0)
    shifts = [(100, 40), (0, 0)]
    full = (4000, 3000)  # full_w, full_h
    # ---- oracle: non_max_suppression + scale_boxes/scale_coords + astype(int) + shift
    y = oy.decode_head(levels)
    dets, cands = oy.non_max_suppression(y, conf, 0.7, 300, return_candidates=True)
    # ---- kernels
    cand, count = ops.pose_decode(_to_dev(levels, cuda_device, False), conf, cap_per_entry=2048)
    cap = cand.shape[1]
    rows = cand.view(-1, ops.ROW)
    seg_off = torch.arange(B, dtype=torch.int32, device=cuda_device) * cap
    res = ops.merge_segments(rows, seg_off, count, cap, merge_type="NMS", metric="IOU", thr=0.7, cmp_strict=True,
                             precision="fp32", class_agnostic=True, pre_cap=30000, max_keep=300, tie_col=5)
    geo = torch.tensor([[shifts[b][0], shifts[b][1], src_w, src_h,
                         round((g["out_w"] - src_w * g["gain"]) / 2 - 0.1), round((g["out_h"] - src_h * g["gain"]) / 2 - 0.1),
                         full[0], full[1]] for b in range(B)], dtype=torch.int32, device=cuda_device)
    fgeo = torch.tensor([[g["gain"], (g["out_w"] - src_w * g["gain"]) / 2, (g["out_h"] - src_h * g["gain"]) / 2, 0.0]] * B,
                        dtype=torch.float32, device=cuda_device)
    grange = torch.tensor([[0, 1], [1, 2]], dtype=torch.int32, device=cuda_device)
    goff = torch.tensor([0, 400], dtype=torch.int32, device=cuda_device)
    det = torch.zeros((800, ops.ROW), dtype=torch.float32, device=cuda_device)
    out_count = torch.zeros((2,), dtype=torch.int32, device=cuda_device)
    ops.finalize_dets(cand, res["keep"], res["keep_count"], geo, fgeo, grange, goff, det, out_count, 400)
    det, out_count, kc = det.cpu(), out_count.cpu(), res["keep_count"].cpu()
    keep = res["keep"].cpu()
    total = 0
    for b in range(B):
        d = dets[b]
        assert int(kc[b]) == d.shape[0] == int(out_count[b])
        # identical kept candidates, in the same (score-descending) order: compare by anchor index
        x, keep_ref = cands[b]
        sc_all = y[b, 4]
        anchors_ref = torch.nonzero(sc_all > conf).flatten()[keep_ref]
        got_rows = rows.cpu()[keep[b * cap: b * cap + int(kc[b])].long()]
        assert got_rows[:, 5].view(torch.int32).tolist() == anchors_ref.tolist()
        boxes = oy.scale_boxes((g["out_h"], g["out_w"]), d[:, :4].clone(), (src_h, src_w))
        kpts = oy.scale_coords((g["out_h"], g["out_w"]), d[:, 6:].reshape(-1, 5, 3).clone(), (src_h, src_w))
        ib = boxes.numpy().astype(int)
        exp = ib + np.array([shifts[b][0], shifts[b][1]] * 2)
        gotb = det[int(goff[b]): int(goff[b]) + d.shape[0]]
        # int truncation may flip where the oracle's float sits within 1e-4 of an integer; everything else is exact
        fl = boxes.numpy()
        safe = np.abs(fl - np.round(fl)) > 1e-4
        assert np.array_equal(gotb[:, :4].numpy().astype(int)[safe], exp[safe])
        assert np.abs(gotb[:, :4].numpy() - exp).max() <= 1
        kexp = kpts.numpy().copy()
        kexp[..., 0] += shifts[b][0]
        kexp[..., 1] += shifts[b][1]
        kg = gotb[:, 6:21].numpy().reshape(-1, 5, 3)
        assert px_close(kg[..., :2], kexp[..., :2])
        assert np.abs(kg[..., 2] - kexp[..., 2]).max() <= SCORE_TOL
        assert np.abs(gotb[:, 4].numpy() - d[:, 4].numpy()).max() <= SCORE_TOL
        total += d.shape[0]
    assert total > 10


@pytest.mark.parametrize("conf", [0.5, 0.25, 0.01, 0.001, 0.9, 0.999])
def test_logit_gate_equals_the_sigmoid_gate(cuda_device, conf):
    """Pass 1 gates on `logit >= x_gate` with x_gate found on the device; that must select exactly the anchors whose
    kernel-side score sigmoid_rn(logit) exceeds conf — for EVERY fp16 logit and a dense sweep of fp32 logits around the gate."""
    import fsd_b200.ops as ops

    def survivors_and_scores(cls_vals, c):
        n = cls_vals.numel()
        side = int(np.ceil(np.sqrt(n / 1.0)))
        h = w = side
        pad = torch.full((h * w - n,), -1e4, dtype=cls_vals.dtype)
        cls0 = torch.cat([cls_vals, pad]).view(1, 1, h, w)
        levels = [(torch.zeros((1, 64, h, w), dtype=cls_vals.dtype), cls0, torch.zeros((1, 15, h, w), dtype=cls_vals.dtype))]
        for _ in range(2):
            levels.append((torch.zeros((1, 64, 1, 1), dtype=cls_vals.dtype), torch.full((1, 1, 1, 1), -1e4, dtype=cls_vals.dtype),
                           torch.zeros((1, 15, 1, 1), dtype=cls_vals.dtype)))
        cand, count = ops.pose_decode(_to_dev(levels, cuda_device, False), c, cap_per_entry=h * w + 2)
        k = int(count[0])
        rows = cand[0, :k].cpu()
        return rows[:, 5].view(torch.int32).long(), rows[:, 4]

    half_all = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.float16)
    half_all = half_all[torch.isfinite(half_all.float())]
    # centre of the fp32 sweep: the gate as numpy's float32 arithmetic sees it (bisection over the ordered bit patterns; for
    # conf 0.5 it is ~3e-8, millions of ulps above 0.0) — the device may land a few ulps away, the sweep covers +-3000
    def sig32(x):
        e = np.float32(np.exp(-np.float64(x)))
        return np.float32(1.0) / (np.float32(1.0) + e)

    def key(x):
        u = int(np.float32(x).view(np.uint32))
        return (~u & 0xffffffff) if u & 0x80000000 else (u | 0x80000000)

    def unkey(k):
        return np.uint32((k & 0x7fffffff) if k & 0x80000000 else (~k & 0xffffffff)).view(np.float32)

    lo_k, hi_k = key(np.float32(-100.0)), key(np.float32(100.0))
    while hi_k - lo_k > 1:
        mid = (lo_k + hi_k) // 2
        if sig32(unkey(mid)) > np.float32(conf):
            hi_k = mid
        else:
            lo_k = mid
    around = [unkey(k) for k in range(hi_k - 3000, hi_k + 3000)]
    for vals in (half_all, torch.tensor(np.array(around, dtype=np.float32))):
        all_idx, all_sc = survivors_and_scores(vals, -1.0)          # conf -1: every anchor passes, scores of all logits
        score = torch.empty(vals.numel())
        keep = all_idx < vals.numel()
        score[all_idx[keep]] = all_sc[keep]
        got_idx, got_sc = survivors_and_scores(vals, conf)
        got = set(got_idx[got_idx < vals.numel()].tolist())
        want = set(torch.nonzero(score > conf).flatten().tolist())
        assert got == want, (len(got), len(want), sorted(got ^ want)[:5])
        assert len(want) > 0 and len(want) < vals.numel()
