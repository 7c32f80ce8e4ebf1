"""Drop-in replacement for the reference module `model/IFNet_m.py` (arbitrary-timestep ECCV-paper IFNet)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _pkg import pkg as _pkg  # noqa: E402

IFNet_m = _pkg.eccv.IFNet_m
IFBlock = _pkg.eccv.IFBlock
Contextnet = _pkg.eccv.Contextnet
Unet = _pkg.eccv.Unet
warp = _pkg.warp
