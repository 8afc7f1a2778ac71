#!/usr/bin/env python
"""Golden vectors for `docs sahi/retinaface_sahi.py` (BASELINE config 3's named file), produced by the
reference's OWN class, imported unmodified through the shims of make_golden.py (insightface.app.FaceAnalysis is the
deterministic fake detector).  The fixture pins the class's behaviour INCLUDING its two quirks under the vendored
`docs sahi/base.py:162-189`:
  * `_create_object_prediction_list_from_original_predictions` RETURNS its list instead of storing it in
    `_object_prediction_list_per_image` (:187-261), so `convert_original_predictions` leaves `object_prediction_list` empty
    and `get_sliced_prediction` finds nothing;
  * the returned objects carry boxes that are ALREADY shifted (:227-230) together with `shift_amount` (:252), so
    `get_shifted_object_prediction()` would shift them twice.

    python tests/golden/make_golden_retinaface.py      # writes tests/golden/retinaface_sahi_outputs.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

import fake_detectors as fd  # noqa: E402
import make_golden as mg  # noqa: E402


def ops_json(preds):
    return [{"bbox": [int(v) for v in p.bbox.to_xyxy()], "shift": [int(v) for v in p.bbox.shift_amount],
             "score": float(p.score.value), "category": [int(p.category.id), p.category.name],
             "shifted": [int(v) for v in p.get_shifted_object_prediction().bbox.to_xyxy()]} for p in preds]


def main():
    predict, pred, base, *_ = mg.load_all()
    rf = mg.load_reference("ref_retinaface_sahi", "docs sahi/retinaface_sahi.py")
    golden = {"ctor": [], "scenes": {}}
    for device, ctx, size in (("cpu", None, 640), ("cpu", 3, "800"), ("cpu", None, None), ("cpu", None, -5)):
        m = rf.RetinaFaceSAHI(confidence_threshold=0.45, device=device, ctx_id=ctx, image_size=size)
        golden["ctor"].append({"device": device, "ctx_id_arg": ctx, "image_size_arg": size, "ctx_id": m.ctx_id,
                               "image_size": m.image_size, "det_size": list(m._resolved_det_size()),
                               "names": m.category_names, "has_mask": m.has_mask, "model_name": m.model_name,
                               "original_predictions_before": list(m.original_predictions)})
    for name, H, W, nf, seed, sl, conf in (("crowd_a", 540, 960, 60, 11, 320, 0.45), ("crowd_b", 384, 512, 30, 12, 256, 0.5)):
        img = fd.coordinate_image(H, W)
        fd.FakeFaceAnalysis.faces = fd.synthetic_faces(H, W, nf, seed)
        m = rf.RetinaFaceSAHI(confidence_threshold=conf, device="cpu", image_size=640)
        window = np.ascontiguousarray(img[64:64 + sl, 128:128 + sl])
        direct = m.perform_inference(window)                       # returns ObjectPredictions (window coordinates)
        n_raw = len(m.original_predictions)
        made = m._create_object_prediction_list_from_original_predictions([128, 64], [H, W])
        made_default = m._create_object_prediction_list_from_original_predictions(None, None)
        m.convert_original_predictions(shift_amount=[128, 64], full_shape=[H, W])
        after_convert = len(m.object_prediction_list)
        res = predict.get_sliced_prediction(img, m, slice_height=sl, slice_width=sl, overlap_height_ratio=0.2,
                                            overlap_width_ratio=0.2, postprocess_type="NMS", postprocess_match_metric="IOU",
                                            postprocess_match_threshold=0.5, verbose=0)
        f32 = m.perform_inference(window.astype(np.float32) / 255.0)  # float input is rescaled and clipped (:104-105)
        golden["scenes"][name] = {"params": [H, W, nf, seed, sl, conf], "direct": ops_json(direct), "n_raw": n_raw,
                                  "created": ops_json(made), "created_default": ops_json(made_default),
                                  "after_convert": after_convert, "sliced": ops_json(res.object_prediction_list),
                                  "float_input": ops_json(f32), "bad_shape": len(m.perform_inference(np.zeros((8, 8), np.uint8))),
                                  "empty": len(m.perform_inference(np.zeros((0, 0, 3), np.uint8)))}
    out = os.path.join(HERE, "retinaface_sahi_outputs.json")
    with open(out, "w") as f:
        json.dump(golden, f, indent=0)
    print("wrote", out, os.path.getsize(out), {k: (len(v["direct"]), len(v["created"]), len(v["sliced"])) for k, v in golden["scenes"].items()})


if __name__ == "__main__":
    main()
