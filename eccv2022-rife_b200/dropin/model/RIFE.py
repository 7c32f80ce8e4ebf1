"""Drop-in replacement for the reference module `model/RIFE.py` (ECCV-paper model facade)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _pkg import pkg as _pkg  # noqa: E402

Model = _pkg.eccv.Model
IFNet = _pkg.eccv.IFNet
warp = _pkg.warp
device = _pkg.eccv.device
