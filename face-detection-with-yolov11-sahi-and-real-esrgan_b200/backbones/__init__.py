"""The conv backbones that stay PyTorch on this path (BASELINE.json north_star): YOLO11n-pose and RRDBNet."""
from .yolo11_pose import YOLO11Pose, build_yolo11n_pose  # noqa: F401
from .rrdbnet import RRDBNet  # noqa: F401
