"""`bbox` import shim: the Cython module of the external WiderFace-Evaluation repo that
eval/eval_official_widerface.py:20-33 imports (`from bbox import bbox_overlaps`), served by the GPU kernel."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from fsd_b200.widerface_eval import bbox_overlaps  # noqa: E402,F401
