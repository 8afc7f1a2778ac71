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
