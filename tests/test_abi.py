"""The C-ABI library loads on a CPU-only box and exports every symbol include/rife_b200.h declares
(no compute calls: those need a GPU)."""
import ctypes as C
import os
import re

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rife_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(rife_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_exported(pkg):
    lib = pkg._lib.load()
    decl = _declared_symbols()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/rife_b200.h but not exported"
    assert sorted(pkg._lib.SYMBOLS) == decl, "eccv2022-rife_b200/_lib.py SYMBOLS out of sync with the header"
    assert lib.rife_abi_version() == 1


def test_error_paths_without_gpu(pkg):
    lib = pkg._lib.load()
    launches = lib.rife_launch_count()
    # shape validation is pure host logic: H, W must be multiples of 64/scale
    assert lib.rife_plan_workspace_bytes(1, 3, 736, 1280, 1.0, 0) == 0
    assert b"multiples of 64" in lib.rife_last_error()
    assert lib.rife_plan_workspace_bytes(1, 3, 1088, 1920, 0.5, 0) == 0          # needs multiples of 128
    assert lib.rife_plan_workspace_bytes(1, 5, 64, 64, 1.0, 0) == 0              # C must be 3 or 4
    assert lib.rife_plan_workspace_bytes(1, 3, 64, 64, 3.0, 0) == 0              # scale must be 0.25 / 0.5 / 1 / 2 / 4
    assert lib.rife_plan_workspace_bytes(1, 3, 32, 96, 2.0, 0) > 0               # scale 2: multiples of 32 suffice
    n32 = lib.rife_plan_workspace_bytes(2, 3, 768, 1280, 1.0, 0)
    n16 = lib.rife_plan_workspace_bytes(2, 3, 768, 1280, 1.0, 1)
    assert n32 > n16 > 0
    # null arguments are rejected before any CUDA call
    assert lib.rife_weights_create(None, 0, None, None, None, 0) == -1
    assert lib.rife_plan_set_profiling(None, 1) == -1
    assert lib.rife_plan_set_chunk(None, 4) == -1
    assert lib.rife_plan_workspace_bytes_ex(3, 3, 1, 3, 64, 64, 1.0, 1) > lib.rife_plan_workspace_bytes(3, 3, 64, 64, 1.0, 1)
    assert lib.rife_scene_ssim(None, None, 0, 1, 3, 8, 8, 32, None, None) == -1
    assert lib.rife_launch_count() == launches                                   # none of the above launched anything
