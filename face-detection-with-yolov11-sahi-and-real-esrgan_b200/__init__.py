"""fsd_b200 — B200-native sliced face-detection hot path (SAHI slice -> batch -> detect -> shift -> merge, plus
Real-ESRGAN tile crop/stitch) behind the reference's own Python API.

Sub-modules:
  _cabi      ctypes binding of include/fsd_b200.h (libfsd_b200.so, hand-written sm_100a kernels)
  ops        torch-tensor front ends of the kernels (device memory + streams only; no torch math)
  sahi_api   drop-in mirror of sahi.predict / sahi.prediction / sahi.annotation / sahi.slicing / sahi.models.base
  plugins    YOLOv11PoseDetectionModel / InsightFaceDetectionModel mirrors (utils/*_wrapper.py)
  enhancer   FaceEnhancer / RealESRGANer mirrors (utils/enhancer.py)
  backbones  the PyTorch conv backbones that stay PyTorch (YOLO11n-pose, RRDBNet)
  shard      image-index data parallelism + the one all-gather of detections
"""
__version__ = "0.1.0"


def install_compat() -> str:
    """Put compat/ (the `sahi` / `ultralytics` / `realesrgan` / `basicsr` / `bbox` import shims) first on sys.path so the
    reference's scripts run unmodified on top of this package.  Returns the directory."""
    import os
    import sys

    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")
    if d not in sys.path:
        sys.path.insert(0, d)
    return d
