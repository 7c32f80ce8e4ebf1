"""Scene-change / static-frame detector of the reference's video loop (inference_video.py:226-228,242-243):
`ssim_matlab(F.interpolate(I0, (32, 32))[:, :3], F.interpolate(I1, (32, 32))[:, :3])` as one fused CUDA kernel
(csrc/scene_ssim.cu).  The reference compares the value with 0.996 (static frame: skip) and 0.2 (scene cut: repeat
the frame instead of interpolating)."""
from __future__ import annotations

import torch

from . import _lib

STATIC_THRESHOLD = 0.996      # inference_video.py:231
SCENE_CUT_THRESHOLD = 0.2     # inference_video.py:246


def scene_ssim(img0, img1, pad_multiple=32):
    """img0 / img1: (B,C,H,W) float32 / float16 frames in [0,1] or (B,H,W,C) uint8 frames on the GPU, C >= 3.
    `pad_multiple`: the frames are taken as zero-padded right / bottom to this multiple before the 32x32 resize, as
    inference_video.py:199-203 pads them (max(32, int(32 / scale))); 1 = no padding.
    Returns a (B,) float32 CUDA tensor -- one SSIM per pair (the reference averages over its batch of one)."""
    if not img0.is_cuda or not img1.is_cuda:
        raise RuntimeError("eccv2022-rife_b200: CUDA tensors required -- no CPU fallback")
    if img0.shape != img1.shape or img0.dtype != img1.dtype:
        raise ValueError("frames must have the same shape and dtype")
    if img0.dtype not in (torch.float32, torch.float16, torch.uint8):
        raise TypeError("float32 / float16 NCHW or uint8 HWC frames required")
    if img0.dtype == torch.uint8:
        b, h, w, c = img0.shape
    else:
        b, c, h, w = img0.shape
    io = {torch.float32: _lib.RIFE_F32, torch.float16: _lib.RIFE_F16, torch.uint8: _lib.RIFE_U8}[img0.dtype]
    img0, img1 = img0.contiguous(), img1.contiguous()
    lib = _lib.load()
    with torch.cuda.device(img0.device):
        out = torch.empty(b, dtype=torch.float32, device=img0.device)
        _lib.check(lib.rife_scene_ssim(img0.data_ptr(), img1.data_ptr(), io, b, c, h, w, int(pad_multiple), out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream))
    return out
