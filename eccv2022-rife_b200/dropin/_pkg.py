"""Locates and imports the `eccv2022-rife_b200` package (its directory name is not a Python identifier)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
pkg = importlib.import_module("eccv2022-rife_b200")
