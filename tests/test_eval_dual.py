"""Secondary evaluator: the oracle restatement and the product module against golden vectors produced by
the reference's own `eval/eval_dual.py` methods (tests/golden/make_golden_eval_dual.py), plus oracle-vs-product on fresh
random cases.  Exact equality: everything is float64 arithmetic in the same order."""
import json
import os

import numpy as np
import pytest

from oracle import eval_dual as oe

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_dual_outputs.json")))


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        if isinstance(a[k], float) or isinstance(b[k], float):
            assert a[k] == pytest.approx(b[k], abs=0, rel=1e-15), k
        else:
            assert a[k] == b[k], k


def test_oracle_matches_reference_iou_and_ap():
    for c in GOLDEN["iou_pairs"]:
        assert oe.calculate_iou(c["a"], c["b"]) == c["iou"]
    for c in GOLDEN["ap_cases"]:
        assert oe.calculate_average_precision([dict(d) for d in c["detections"]], c["total_gt"]) == pytest.approx(c["ap"], abs=0, rel=1e-15)


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"seed{c['seed']}")
def test_oracle_and_product_match_reference_sets(case):
    import fsd_b200.eval_dual as pe

    for mod in (oe, pe):
        sub, diff, summary = mod.evaluate_all(case["gt"], case["predictions"])
        for got, want in zip(sub, case["subcategory"]):
            _same(got, want)
        for got, want in zip(diff, case["difficulty"]):
            _same(got, want)
        _same({k: float(v) for k, v in summary.items()}, case["summary"])
    for cat, want in case["difficulty_of"].items():
        assert oe.map_subcategory_to_difficulty(cat) == want


def test_product_matches_oracle_on_random_cases_and_matrix_iou():
    import fsd_b200.eval_dual as pe

    rng = np.random.default_rng(5)
    a = rng.integers(0, 50, (30, 4)).astype(float)
    b = rng.integers(0, 50, (20, 4)).astype(float)
    m = pe.iou_xywh_matrix(a, b)
    for i in range(30):
        for j in range(20):
            assert m[i, j] == oe.calculate_iou(list(a[i]), list(b[j]))
    for n, total in ((0, 3), (5, 0), (30, 11), (200, 150)):
        conf = rng.uniform(0, 1, n).round(2)
        tp = rng.random(n) < 0.5
        dets = [{"confidence": float(c), "is_tp": bool(t)} for c, t in zip(conf, tp)]
        assert pe.average_precision_11pt(conf, tp, total) == pytest.approx(oe.calculate_average_precision(dets, total), abs=0, rel=1e-15)
    # thresholds other than the defaults, incl. iou_threshold = 0 (a zero-IoU best match is never a true positive)
    case = GOLDEN["cases"][0]
    for thr, conf_thr in ((0.3, 0.1), (0.75, 0.6), (0.0, 0.25)):
        for c in oe.SUBCATEGORIES[:3]:
            _same(pe.evaluate_single_set(case["gt"], case["predictions"], c, [c], thr, conf_thr),
                  oe.evaluate_single_set(case["gt"], case["predictions"], c, [c], thr, conf_thr))


def test_predictions_from_results():
    import fsd_b200.eval_dual as pe
    from fsd_b200.sahi_api.prediction import ObjectPrediction, PredictionResult

    preds = [ObjectPrediction(bbox=[10, 20, 50, 80], category_id=0, category_name="face", score=0.9),
             ObjectPrediction(bbox=[0, 0, 8, 8], category_id=0, category_name="face", score=0.3)]
    res = PredictionResult(object_prediction_list=preds, image=np.zeros((100, 100, 3), np.uint8))
    out = pe.predictions_from_results([res], ["a.jpg"])
    assert out == {"a.jpg": [{"bbox": [10, 20, 40, 60], "confidence": 0.9}, {"bbox": [0, 0, 8, 8], "confidence": 0.3}]}
    assert pe.predictions_from_results([res], ["a.jpg"], scale=2.0)["a.jpg"][0]["bbox"] == [5.0, 10.0, 20.0, 30.0]
