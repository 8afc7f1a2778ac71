"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy / torch-CPU / cv2) of the reference's sliced-inference hot path
(ihsanhadi57/Face-Detection-With-YOLOv11-SAHI-and-Real-ESRGAN): SAHI slice -> detect -> shift -> merge behind
`docs sahi/predict.py`, `utils/yolo_wrapper.py`, and the Real-ESRGAN tile crop/stitch behind `utils/enhancer.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import
this package, and only as the checker or the CPU baseline — never as the product path.  The product
(`fsd_b200`) never imports it and raises when its CUDA library is missing.

PARITY PINNING.  The reference has no tests, fixtures or golden vectors, and the arithmetic of this path
lives in third-party packages that are neither vendored under /root/reference nor installable here
(sahi==0.11.34, shapely==2.0.7, ultralytics (unpinned), realesrgan==0.3.0, basicsr==1.4.2).  What IS pinned:
  * orchestration, plugin conversion and data classes: the reference's own `docs sahi/{predict,prediction,
    base}.py`, `utils/yolo_wrapper.py`, `utils/insightface_wrapper.py` and `utils/enhancer.py` are imported
    UNMODIFIED by `tests/golden/make_golden.py` (through thin `sahi`/`ultralytics`/`realesrgan` import shims)
    and their outputs are committed as fixtures under tests/golden/ — the oracle must reproduce them;
  * the secondary evaluator: `eval/eval_dual.py` is imported unmodified by `tests/golden/make_golden_eval_dual.py` and the
    results of its IoU / 11-point AP / matching methods are committed (tests/golden/eval_dual_outputs.json);
  * cv2.resize / copyMakeBorder (letterbox) and torchvision.ops.nms: the real libraries are in this image and
    the integer restatement in `oracle/letterbox.py` / `oracle/yolo_head.py` is checked against them.
What is "parity unpinned": the published algorithms of sahi.slicing / sahi.postprocess / sahi.annotation,
ultralytics' head decode + NMS + rescale, and realesrgan's tile_process are restated from their documented
behaviour with every ambiguity resolved by a stated default.
"""
