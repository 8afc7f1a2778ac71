"""sahi.postprocess.combine mirror: the four Postprocess classes selected at docs sahi/predict.py:44-49.

`__call__(list[ObjectPrediction]) -> list[ObjectPrediction]` keeps the upstream contract, but the match /
suppress / merge arithmetic runs in Kernel 3 on the GPU (one segment = the given list).  There is no CPU path:
without the CUDA library the call raises."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from .. import ops
from .prediction import ObjectPrediction


def predictions_to_rows(object_predictions) -> np.ndarray:
    """ObjectPredictionList.totensor(): rows [x1,y1,x2,y2,score,category_id] float32."""
    rows = np.empty((len(object_predictions), 6), dtype=np.float32)
    for i, op in enumerate(object_predictions):
        b = op.bbox
        rows[i, 0], rows[i, 1], rows[i, 2], rows[i, 3] = b.minx, b.miny, b.maxx, b.maxy
        rows[i, 4] = op.score.value
        rows[i, 5] = op.category.id
    return rows


class PostprocessPredictions:
    """Base: holds the match configuration (upstream defaults: 0.5 / IOU / class-agnostic)."""

    merge_type = None

    def __init__(self, match_threshold: float = 0.5, match_metric: str = "IOU", class_agnostic: bool = True,
                 tie_rule: str = "box_lex"):
        if match_metric not in ("IOU", "IOS"):
            raise ValueError(f"'match_metric' should be one of ['IOU', 'IOS'] but given as {match_metric}")
        self.match_threshold = match_threshold
        self.match_metric = match_metric
        self.class_agnostic = class_agnostic
        # equal scores: "box_lex" = sahi 0.11.34's lexicographic rule (SURVEY A.2.4 variant N), "index" = plain greedy order
        self.tie_rule = tie_rule
        self.device = None  # set by get_sliced_prediction; defaults to the current CUDA device

    def _run(self, object_predictions):
        if not torch.cuda.is_available():
            raise ops._cabi.FsdError("fsd_b200 postprocess needs a CUDA device: there is no CPU fallback")
        dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        rows = torch.from_numpy(predictions_to_rows(object_predictions)).to(dev)
        n = rows.shape[0]
        cats = rows[:, 5].to(torch.int32).contiguous()
        res = ops.merge_segments(rows, torch.zeros(1, dtype=torch.int32, device=dev), None, n,
                                 merge_type=self.merge_type, metric=self.match_metric, thr=self.match_threshold,
                                 cmp_strict=False, precision="fp64", class_agnostic=self.class_agnostic, cats=cats,
                                 want_parent=False, tie_rule=self.tie_rule)
        k = int(res["keep_count"][0])
        keep = res["keep"][:k].cpu().numpy()
        order = np.arange(k)
        if not self.class_agnostic and self.merge_type != "NMS":
            # upstream batched_* variants emit category by category (ascending id), score-descending inside
            order = np.argsort(rows[:, 5].cpu().numpy()[keep], kind="stable")
        return keep, order, res, k

    def __call__(self, object_predictions: List[ObjectPrediction]):
        raise NotImplementedError()


class NMSPostprocess(PostprocessPredictions):
    merge_type = "NMS"

    def __call__(self, object_predictions):
        if len(object_predictions) == 0:
            return []
        keep, order, _, _ = self._run(object_predictions)
        return [object_predictions[int(i)] for i in keep[order]]


class _MergingPostprocess(PostprocessPredictions):
    def __call__(self, object_predictions):
        if len(object_predictions) == 0:
            return []
        keep, order, res, k = self._run(object_predictions)
        boxes = res["boxes"][:k].cpu().numpy()
        cats = res["cats"][:k].cpu().numpy()
        names = {}
        for op in object_predictions:
            names.setdefault(op.category.id, op.category.name)
        out = []
        for j in order:
            src = object_predictions[int(keep[j])]
            b = boxes[j]
            src_box = src.bbox.to_xyxy()
            if all(float(b[c]) == float(src_box[c]) for c in range(4)):
                merged_box = src_box  # nothing merged: keep the original coordinate objects
            else:
                integral = all(float(v).is_integer() for v in b)
                merged_box = [int(v) for v in b] if integral else [float(v) for v in b]
            cid = int(cats[j])
            merged = ObjectPrediction(bbox=merged_box, score=src.score.value, category_id=cid,
                                      category_name=names.get(cid), segmentation=None,
                                      shift_amount=list(src.bbox.shift_amount), full_shape=None)
            if hasattr(src, "keypoints"):
                merged.keypoints = src.keypoints
            out.append(merged)
        return out


class GreedyNMMPostprocess(_MergingPostprocess):
    merge_type = "GREEDYNMM"


class NMMPostprocess(_MergingPostprocess):
    merge_type = "NMM"


class LSNMSPostprocess(PostprocessPredictions):
    def __call__(self, object_predictions):
        raise NotImplementedError("LSNMS depends on the external `lsnms` package and is never selected by the reference; "
                                  "use NMS, GREEDYNMM or NMM")
