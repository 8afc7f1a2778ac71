"""`sahi` import shim -> fsd_b200.sahi_api (see compat/README.md)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import fsd_b200  # noqa: E402,F401

__version__ = "0.11.34+fsd_b200"
