"""Drop-in replacement for the reference module `model/IFNet_2R.py` (with `model/refine_2R.py`): the stride variant of
the ECCV-paper IFNet.  The reference names the class `IFNet` in that module too."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _pkg import pkg as _pkg  # noqa: E402

IFNet = _pkg.eccv.IFNet_2R
IFBlock = _pkg.eccv.IFBlock
Contextnet = _pkg.eccv.Contextnet
Unet = _pkg.eccv.Unet
warp = _pkg.warp
