"""Drop-in replacement for the reference module `train_log/IFNet_HDv3.py` (same import path and names)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _pkg import pkg as _pkg  # noqa: E402

IFNet = _pkg.IFNet
IFBlock = _pkg.IFBlock
Head = _pkg.Head
ResConv = _pkg.ResConv
warp = _pkg.warp
device = _pkg.model.device
