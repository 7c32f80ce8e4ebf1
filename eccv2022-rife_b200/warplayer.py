"""Drop-in for the reference's `model/warplayer.py:8-22` `warp(tenInput, tenFlow)` on CUDA tensors."""
from __future__ import annotations

import torch

from . import _lib

_grid_cache = {}


def base_grids(h, w, device):
    """The base grids torch.linspace(-1, 1, W) / (-1, 1, H) of model/warplayer.py:11-14, computed with ATen's CPU
    linspace kernel (float32) and uploaded once per (device, H, W).  The reference evaluates the same expression on its
    module-level `device`, i.e. with the CUDA linspace kernel when a GPU is present; the two kernels may differ by one
    ulp at some entries (as ATen's vectorised CPU kernel differs from the scalar formula), which
    moves a warped pixel by ~1e-5.  The goldens are recorded from the reference on CPU, so the CPU grid is what is
    pinned; tests/test_gpu_ops.py::test_base_grid_vs_cuda_linspace bounds the difference to the CUDA kernel.
    Returns (host_x, host_y, dev_x, dev_y)."""
    k = (str(device), h, w)
    if k not in _grid_cache:
        gx = torch.linspace(-1.0, 1.0, w, dtype=torch.float32)
        gy = torch.linspace(-1.0, 1.0, h, dtype=torch.float32)
        _grid_cache[k] = (gx, gy, gx.to(device), gy.to(device))
    return _grid_cache[k]


def warp(tenInput, tenFlow):
    if not (tenInput.is_cuda and tenFlow.is_cuda):
        raise RuntimeError("eccv2022-rife_b200.warp: CUDA tensors required (the B200 path has no CPU fallback)")
    if tenInput.dtype not in (torch.float32, torch.float16) or tenFlow.dtype != tenInput.dtype:
        raise TypeError("warp: float32 or float16 tensors of matching dtype required")
    b, c, h, w = tenInput.shape
    if tuple(tenFlow.shape) != (b, 2, h, w):
        raise ValueError(f"warp: flow must be ({b}, 2, {h}, {w}), got {tuple(tenFlow.shape)}")
    lib = _lib.load()
    x = tenInput.contiguous()
    fl = tenFlow.contiguous()
    out = torch.empty_like(x)
    _, _, gx, gy = base_grids(h, w, x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.rife_warp(x.data_ptr(), fl.data_ptr(), out.data_ptr(), b, c, h, w,
                                 _lib.RIFE_F32 if x.dtype == torch.float32 else _lib.RIFE_F16,
                                 gx.data_ptr(), gy.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out
