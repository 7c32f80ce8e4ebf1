"""Multi-pass interpolation helpers of the reference's application layer, on top of the batched backend.

Same names, arguments and results as `rife_app/utils/interpolation.py:4-95`; the difference is how the work is issued:
the reference calls `model.inference` once per adjacent pair of every refinement stage (2**n - 1 calls), here every
stage is ONE batched call over all of its pairs (n calls), which is what fills the SMs at small frame sizes."""
from __future__ import annotations

import torch


def _rounds(first, last, n_exp, model, scale):
    frames = [first, last]
    b = first.shape[0]
    for _ in range(int(n_exp)):
        left = torch.cat(frames[:-1], 0)
        right = torch.cat(frames[1:], 0)
        mids = model.inference(left, right, scale=scale)           # timestep stays 0.5 (interpolation.py:33-34)
        nxt = []
        for j in range(len(frames) - 1):
            nxt.append(frames[j])
            nxt.append(mids[j * b:(j + 1) * b])
        nxt.append(frames[-1])
        frames = nxt
    return frames


def generate_interpolated_frames(img0_tensor, img1_tensor, exp_value, model):
    """rife_app/utils/interpolation.py:4-19 (the loop of inference_img.py:95-103): [img0, ..., img1], 2**exp + 1 frames."""
    return _rounds(img0_tensor, img1_tensor, exp_value, model, 1.0)


def progressive_interpolate_frames(I0, I1, n_exp, model, model_scale_factor):
    """rife_app/utils/interpolation.py:28-95: the 2**n_exp - 1 in-between frames in temporal order, every one of them the
    timestep-0.5 interpolation of its two neighbours of the previous stage."""
    if n_exp == 0:
        return []
    return _rounds(I0, I1, n_exp, model, model_scale_factor)[1:-1]


def recursive_interpolate_video_frames(I0, I1, n_exp, model, model_scale_factor):
    """rife_app/utils/interpolation.py:21-26 (deprecated alias kept by the reference)."""
    return progressive_interpolate_frames(I0, I1, n_exp, model, model_scale_factor)
