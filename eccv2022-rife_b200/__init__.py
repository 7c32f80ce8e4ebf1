"""eccv2022-rife_b200: B200-native (sm_100a) implementation of the RIFE HDv3 (v4.25) IFNet inference hot path.

Import with `importlib.import_module("eccv2022-rife_b200")` (the directory name is not a Python identifier), or
put `eccv2022-rife_b200/dropin` in front of the reference on sys.path and keep the reference's own imports
(`from train_log.RIFE_HDv3 import Model`, `from model.warplayer import warp`).
"""
from . import _lib
from ._lib import RifeError
from .ifnet import IFNet, IFBlock, Head, ResConv
from .model import Model
from .warplayer import warp
from .sharding import shard_range, run_sharded
from . import eccv
from . import scene
from .scene import scene_ssim

__all__ = ["Model", "IFNet", "IFBlock", "Head", "ResConv", "warp", "RifeError", "shard_range", "run_sharded"]
