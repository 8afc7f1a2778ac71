"""Oracle for the cross-slice merge (test infrastructure, see oracle/__init__).

Restates [EXT sahi==0.11.34] sahi.postprocess.combine / sahi.postprocess.utils as recorded in SURVEY.md
App. A.2 — selected by the reference at docs sahi/predict.py:44-49,250-259 and run at :297,:319.
Parity unpinned (upstream source and its shapely/GEOS dependency are absent).  Defaults chosen where upstream
generations differ: float64 metric, match when metric >= threshold, zero-area boxes never match,
visiting order = score descending with ties visited in original index order, equal scores handled by variant N's
lexicographic box rule (`tie_rule="box_lex"`, the default; `"index"` is the plain greedy loop), candidates of a keep listed
in rank order (GREEDYNMM) / in append order of the transitive walk (NMM).
"""
from __future__ import annotations

import numpy as np

from .annotation import ObjectPrediction


# ---- tensor form: rows [x1,y1,x2,y2,score,category] float32 (ObjectPredictionList.totensor) -------------
def to_array(object_predictions) -> np.ndarray:
    arr = np.zeros((len(object_predictions), 6), dtype=np.float32)
    for i, p in enumerate(object_predictions):
        arr[i, :4] = p.bbox.to_xyxy()
        arr[i, 4] = p.score.value
        arr[i, 5] = p.category.id
    return arr


def rank_order(scores: np.ndarray) -> np.ndarray:
    """score descending, ties -> lower index first (stable)."""
    return np.argsort(-scores.astype(np.float64), kind="stable")


def metric_row(boxes: np.ndarray, i: int, js: np.ndarray, match_metric: str, dtype=np.float64) -> np.ndarray:
    """match metric between box i and boxes js; 0 where the denominator is 0 (shapely-era sahi behaviour)."""
    b = boxes.astype(dtype)
    x1 = np.maximum(b[js, 0], b[i, 0]); y1 = np.maximum(b[js, 1], b[i, 1])
    x2 = np.minimum(b[js, 2], b[i, 2]); y2 = np.minimum(b[js, 3], b[i, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    area_i = (b[i, 2] - b[i, 0]) * (b[i, 3] - b[i, 1])
    area_j = (b[js, 2] - b[js, 0]) * (b[js, 3] - b[js, 1])
    if match_metric == "IOU":
        den = area_i + area_j - inter
    elif match_metric == "IOS":
        den = np.minimum(area_j, area_i)
    else:
        raise ValueError(f"unknown match_metric {match_metric}")
    out = np.zeros_like(inter)
    np.divide(inter, den, out=out, where=den > 0)
    return out


TIE_RULE = "box_lex"  # module default: what sahi 0.11.34 is believed to ship (SURVEY A.2.4 variant N); "index" = plain order


def lex_greater(boxes: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """tuple(box) > tuple(ref), row by row (python tuple comparison of the four coordinates)."""
    out = np.zeros(len(boxes), dtype=bool)
    undecided = np.ones(len(boxes), dtype=bool)
    for c in range(4):
        out |= undecided & (boxes[:, c] > ref[c])
        undecided &= boxes[:, c] == ref[c]
    return out


def _candidates(order, r, boxes, scores, suppressed, tie_rule):
    """Indices the current box order[r] is tested against, in rank order.

    tie_rule "index": every not yet suppressed box of lower rank (ties broken by the lower original index) — the greedy
    loop of the pure-torch generations.
    tie_rule "box_lex" (SURVEY A.2.4 variant N, lines 606-617): every not suppressed box whose score is not higher, EXCEPT
    equal-score boxes whose coordinate tuple is lexicographically larger than the current one's.  Two consequences, both
    restated faithfully: (a) an equal-score, lexicographically larger box of lower rank is NOT suppressed by the current
    box, so both can be kept; (b) when that box is visited later it DOES test — and can "suppress" — the earlier equal-score
    keep, which for GREEDYNMM puts an already kept box into the later keep's merge list."""
    cur = order[r]
    if tie_rule == "index":
        cand = order[r + 1:]
        return cand[~suppressed[cand]]
    if tie_rule != "box_lex":
        raise ValueError(f"unknown tie_rule {tie_rule}")
    lo = r
    while lo > 0 and scores[order[lo - 1]] == scores[cur]:
        lo -= 1
    cand = np.concatenate((order[lo:r], order[r + 1:]))
    cand = cand[~suppressed[cand]]
    skip = (scores[cand] == scores[cur]) & lex_greater(boxes[cand], boxes[cur])
    return cand[~skip]


def nms(preds: np.ndarray, match_metric="IOU", match_threshold=0.5, tie_rule=None):
    boxes, scores = preds[:, :4], preds[:, 4]
    tie_rule = tie_rule or TIE_RULE
    order = rank_order(scores)
    suppressed = np.zeros(len(preds), dtype=bool)
    keep = []
    for r, cur in enumerate(order):
        if suppressed[cur]:
            continue
        keep.append(int(cur))
        rest = _candidates(order, r, boxes, scores, suppressed, tie_rule)
        if len(rest):
            m = metric_row(boxes, cur, rest, match_metric) >= match_threshold
            suppressed[rest[m]] = True
    return keep


def greedy_nmm(preds: np.ndarray, match_metric="IOU", match_threshold=0.5, tie_rule=None):
    boxes, scores = preds[:, :4], preds[:, 4]
    tie_rule = tie_rule or TIE_RULE
    order = rank_order(scores)
    suppressed = np.zeros(len(preds), dtype=bool)
    keep_to_merge = {}
    for r, cur in enumerate(order):
        if suppressed[cur]:
            continue
        rest = _candidates(order, r, boxes, scores, suppressed, tie_rule)
        merged = []
        if len(rest):
            m = metric_row(boxes, cur, rest, match_metric) >= match_threshold
            merged = [int(j) for j in rest[m]]
            suppressed[rest[m]] = True
        keep_to_merge[int(cur)] = merged
    return keep_to_merge


def nmm(preds: np.ndarray, match_metric="IOU", match_threshold=0.5, tie_rule=None):
    """Transitive merge (SURVEY A.2.4 `nmm`): every box is visited in rank order; a box already claimed by a keep
    forwards its own unclaimed matches to that keep.  Matches are listed in ascending-score order (`flip`)."""
    boxes, scores = preds[:, :4], preds[:, 4]
    order = rank_order(scores)
    keep_to_merge, merge_to_keep = {}, {}
    for cur in order:
        cur = int(cur)
        others = order[order != cur]
        m = metric_row(boxes, cur, others, match_metric) >= match_threshold
        matched = [int(j) for j in others[m][::-1]]
        if cur not in merge_to_keep:
            keep_to_merge[cur] = []
            for j in matched:
                if j not in merge_to_keep:
                    keep_to_merge[cur].append(j)
                    merge_to_keep[j] = cur
        else:
            k = merge_to_keep[cur]
            for j in matched:
                if j not in keep_to_merge and j not in merge_to_keep:
                    keep_to_merge[k].append(j)
                    merge_to_keep[j] = k
    return keep_to_merge


def _batched(fn_dict_or_list, preds, match_metric, match_threshold, is_dict, **kw):
    cats = preds[:, 5]
    if is_dict:
        out = {}
        for c in np.unique(cats):
            idx = np.where(cats == c)[0]
            sub = fn_dict_or_list(preds[idx], match_metric, match_threshold, **kw)
            for k, lst in sub.items():
                out[int(idx[k])] = [int(idx[j]) for j in lst]
        return out
    mask = np.zeros(len(preds), dtype=bool)
    for c in np.unique(cats):
        idx = np.where(cats == c)[0]
        mask[idx[fn_dict_or_list(preds[idx], match_metric, match_threshold, **kw)]] = True
    keep = np.where(mask)[0]
    return [int(i) for i in keep[rank_order(preds[keep, 4])]]


def batched_nms(preds, match_metric="IOU", match_threshold=0.5, tie_rule=None):
    return _batched(nms, preds, match_metric, match_threshold, False, tie_rule=tie_rule)


def batched_greedy_nmm(preds, match_metric="IOU", match_threshold=0.5, tie_rule=None):
    return _batched(greedy_nmm, preds, match_metric, match_threshold, True, tie_rule=tie_rule)


def batched_nmm(preds, match_metric="IOU", match_threshold=0.5, tie_rule=None):
    return _batched(nmm, preds, match_metric, match_threshold, True)  # variant N's nmm has no tie rule (A.2.4)


# ---- object level: has_match / merge (sahi.postprocess.utils) ----------------------------
def _area(b):
    return (b[2] - b[0]) * (b[3] - b[1])


def _intersection(b1, b2):
    lt = np.maximum(b1[:2], b2[:2])
    rb = np.minimum(b1[2:], b2[2:])
    wh = (rb - lt).clip(min=0)
    return wh[0] * wh[1]


def has_match(p1, p2, match_type="IOU", match_threshold=0.5) -> bool:
    b1, b2 = np.array(p1.bbox.to_xyxy()), np.array(p2.bbox.to_xyxy())
    inter = _intersection(b1, b2)
    with np.errstate(divide="ignore", invalid="ignore"):
        if match_type == "IOU":
            v = inter / (_area(b1) + _area(b2) - inter)
        elif match_type == "IOS":
            v = inter / np.minimum(_area(b1), _area(b2))
        else:
            raise ValueError()
    return bool(v > match_threshold)  # STRICT, unlike the >= of the tensor pass


def merge_object_prediction_pair(p1, p2):
    b1, b2 = np.array(p1.bbox.to_xyxy()), np.array(p2.bbox.to_xyxy())
    box = list(np.concatenate((np.minimum(b1[:2], b2[:2]), np.maximum(b1[2:], b2[2:]))))
    score = max(p1.score.value, p2.score.value)
    cat = p1.category if p1.score.value > p2.score.value else p2.category
    return ObjectPrediction(bbox=box, score=score, category_id=cat.id, category_name=cat.name, segmentation=None,
                            shift_amount=p1.bbox.shift_amount, full_shape=None)


class PostprocessPredictions:
    def __init__(self, match_threshold=0.5, match_metric="IOU", class_agnostic=True, tie_rule=None):
        self.match_threshold, self.match_metric, self.class_agnostic = match_threshold, match_metric, class_agnostic
        self.tie_rule = tie_rule  # None: the module default TIE_RULE

    def __call__(self, object_predictions):
        raise NotImplementedError()


class NMSPostprocess(PostprocessPredictions):
    def __call__(self, object_predictions):
        preds = to_array(object_predictions)
        fn = nms if self.class_agnostic else batched_nms
        keep = fn(preds, self.match_metric, self.match_threshold, tie_rule=self.tie_rule)
        self.last_keep = keep
        return [object_predictions[i] for i in keep]


class _MergePostprocess(PostprocessPredictions):
    _plain = _batched_fn = None

    def __call__(self, object_predictions):
        opl = list(object_predictions)
        preds = to_array(opl)
        fn = type(self)._plain if self.class_agnostic else type(self)._batched_fn
        keep_to_merge = fn(preds, self.match_metric, self.match_threshold, tie_rule=self.tie_rule)
        self.last_keep_to_merge = keep_to_merge
        selected = []
        for keep_ind, merge_list in keep_to_merge.items():
            for merge_ind in merge_list:
                if has_match(opl[keep_ind], opl[merge_ind], self.match_metric, self.match_threshold):
                    opl[keep_ind] = merge_object_prediction_pair(opl[keep_ind], opl[merge_ind])
            selected.append(opl[keep_ind])
        return selected


class GreedyNMMPostprocess(_MergePostprocess):
    _plain, _batched_fn = staticmethod(greedy_nmm), staticmethod(batched_greedy_nmm)


class NMMPostprocess(_MergePostprocess):
    _plain, _batched_fn = staticmethod(nmm), staticmethod(batched_nmm)


class LSNMSPostprocess(PostprocessPredictions):
    def __call__(self, object_predictions):
        raise NotImplementedError("LSNMS needs the external `lsnms` package and no reference caller selects it")


POSTPROCESS_NAME_TO_CLASS = {"GREEDYNMM": GreedyNMMPostprocess, "NMM": NMMPostprocess, "NMS": NMSPostprocess,
                             "LSNMS": LSNMSPostprocess}
