"""(f1) WIDER-FACE AP: oracle restatement sanity (CPU) and GPU-overlap evaluator parity."""
import numpy as np
import pytest

from oracle import widerface_eval as oe


def _scene(rng, n_img=12):
    preds, gts, keeps = [], [], {"easy": [], "medium": [], "hard": []}
    for _ in range(n_img):
        k = int(rng.integers(0, 9))
        gt = np.stack([rng.uniform(0, 900, k), rng.uniform(0, 600, k), rng.uniform(8, 120, k), rng.uniform(8, 150, k)], 1) if k else np.zeros((0, 4))
        det = []
        for g in gt:
            if rng.random() < 0.8:
                det.append([*(g[:2] + rng.normal(0, 2, 2)), *(g[2:] * rng.uniform(0.9, 1.1, 2)), rng.uniform(0.3, 1)])
        for _ in range(int(rng.integers(0, 4))):
            det.append([rng.uniform(0, 900), rng.uniform(0, 600), rng.uniform(8, 80), rng.uniform(8, 80), rng.uniform(0.01, 0.6)])
        det = np.array(sorted(det, key=lambda r: -r[4]), dtype=float).reshape(-1, 5)
        preds.append(det)
        gts.append(gt)
        kl = oe.difficulty_keep_lists(gt)
        for s in keeps:
            keeps[s].append(kl[s])
    return preds, gts, keeps


def test_oracle_ap_sanity():
    gt = [np.array([[10.0, 10, 50, 60], [200.0, 100, 40, 40]])]
    perfect = [np.array([[10.0, 10, 50, 60, 0.9], [200.0, 100, 40, 40, 0.8]])]
    ap, _ = oe.evaluate_setting(perfect, gt, [np.array([1, 2])], thresh_num=100)
    assert ap == pytest.approx(1.0)
    ap, _ = oe.evaluate_setting([np.array([[500.0, 500, 10, 10, 0.9]])], gt, [np.array([1, 2])], thresh_num=100)
    assert ap == 0.0
    assert oe.bbox_overlaps(np.array([[0.0, 0, 9, 9]]), np.array([[0.0, 0, 9, 9], [5.0, 5, 14, 14]])).tolist() == [[1.0, 25 / 175]]


@pytest.mark.gpu
def test_gpu_evaluator_matches_oracle(cuda_device):
    import fsd_b200.widerface_eval as pe

    rng = np.random.default_rng(3)
    preds, gts, keeps = _scene(rng, 20)
    b, q = rng.uniform(0, 500, (37, 4)), rng.uniform(0, 500, (11, 4))
    b[:, 2:] += b[:, :2]
    q[:, 2:] += q[:, :2]
    assert np.array_equal(pe.bbox_overlaps(b, q), oe.bbox_overlaps(b, q))
    for s in ("easy", "medium", "hard"):
        ap_o, curve_o = oe.evaluate_setting(preds, gts, keeps[s], thresh_num=1000)
        ap_p, curve_p = pe.evaluate_setting(preds, gts, keeps[s], thresh_num=1000)
        assert ap_p == ap_o and np.array_equal(curve_p, curve_o)


# ---- pinned to the reference's own evaluator methods (tests/golden/make_golden_widerface_eval.py) ---------------------
import hashlib  # noqa: E402
import json  # noqa: E402
import os  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "widerface_eval_outputs.json")


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def _flat(case, setting):
    preds, gts, keeps = [], [], []
    for ev in case["events"]:
        for im in ev["images"]:
            preds.append(np.array(im["pred"], dtype=float).reshape(-1, 5))
            gts.append(np.array(im["gt"], dtype=float).reshape(-1, 4))
            keeps.append(np.array(im["keep"][setting], dtype=np.int64))
    return preds, gts, keeps


def _check_against_reference_golden(mod, **kw):
    g = json.load(open(GOLDEN))
    for c in g["voc_ap"]:
        assert float(mod.voc_ap(np.array(c["rec"]), np.array(c["prec"]))) == c["ap"]
    case = g["cases"][0]
    images = {im["name"]: im for ev in case["events"] for im in ev["images"]}
    assert len(g["image_eval"]) >= 8
    for rec in g["image_eval"]:  # eval/eval_official_widerface.py:302-377 on every image of case 0 (medium setting)
        im = images[rec["image"]]
        pred, gt = np.array(im["pred"], dtype=float), np.array(im["gt"], dtype=float)
        pr, pl = mod.image_eval(pred.copy(), gt.copy(), np.array(rec["ignore"]), 0.5, **kw)
        assert pr.tolist() == rec["pred_recall"] and pl.tolist() == rec["proposal_list"]
        info = mod.img_pr_info(1000, pred, pl, pr)
        assert _digest(info) == rec["pr_info_sha256"]
    for case in g["cases"]:
        for setting, want in case["settings"].items():
            preds, gts, keeps = _flat(case, setting)
            ap, curve = mod.evaluate_setting(preds, gts, keeps, thresh_num=1000, **kw)
            assert float(ap) == want["ap"], (setting, ap, want["ap"])
            assert _digest(curve[:, 1]) == want["recall_sha256"] and _digest(curve[:, 0]) == want["propose_sha256"]
            if "recall" in want:
                assert curve[:, 1].tolist() == want["recall"] and curve[:, 0].tolist() == want["propose"]


def test_oracle_reproduces_reference_evaluator_golden():
    _check_against_reference_golden(oe)
    g = json.load(open(GOLDEN))
    curve = np.stack([np.arange(1000.0) % 7, np.arange(1000.0) % 5], 1)
    assert _digest(oe.dataset_pr_info(1000, curve, g["dataset_pr_info"]["count_face"])) == g["dataset_pr_info"]["sha256"]


@pytest.mark.gpu
def test_device_evaluator_reproduces_reference_evaluator_golden(cuda_device):
    """The product's evaluator — IoU(+1), greedy matching and the 1000-threshold PR accumulation all in ONE kernel launch
    per setting (fsd_widerface_pr_curve) — against the reference's own methods."""
    import fsd_b200.widerface_eval as pe

    _check_against_reference_golden(pe)
