"""`ultralytics` import shim: only the YOLO front end the reference uses (utils/yolo_wrapper.py:4)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from fsd_b200.yolo import YOLO, Results  # noqa: E402,F401
