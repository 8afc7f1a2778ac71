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
