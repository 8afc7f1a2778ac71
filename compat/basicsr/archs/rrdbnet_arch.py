"""`basicsr.archs.rrdbnet_arch` import shim (utils/enhancer.py:11)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from fsd_b200.backbones.rrdbnet import RRDBNet  # noqa: E402,F401
