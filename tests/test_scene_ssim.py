"""Scene-change statistic of the video loop (inference_video.py:226-228): oracle pinned to values recorded from the
reference's own F.interpolate + ssim_matlab, CUDA kernel against both."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import torch_ref as T


def _nchw(u8):
    return torch.from_numpy(u8.transpose(0, 3, 1, 2).copy()).float() / 255.


def test_scene_ssim_oracle_against_golden():
    g = load_golden("scene_ssim_cases")
    a, b = _nchw(g["img0"]), _nchw(g["img1"])
    for i in range(len(g["ssim"])):
        assert abs(T.scene_ssim(a[i:i + 1], b[i:i + 1], 32) - g["ssim"][i]) < 2e-6, i


@pytest.mark.gpu
def test_scene_ssim_kernel(pkg):
    g = load_golden("scene_ssim_cases")
    u0, u1 = torch.from_numpy(g["img0"]).cuda(), torch.from_numpy(g["img1"]).cuda()
    got = pkg.scene_ssim(u0, u1, 32).cpu().numpy()                      # uint8 HWC frames, batch of 7
    assert got.shape == (7,) and np.abs(got - g["ssim"]).max() < 2e-5   # fp32, separable vs 11^3 summation order
    f0, f1 = _nchw(g["img0"]).cuda(), _nchw(g["img1"]).cuda()
    assert np.abs(pkg.scene_ssim(f0, f1, 32).cpu().numpy() - g["ssim"]).max() < 2e-5
    assert np.abs(pkg.scene_ssim(f0.half(), f1.half(), 32).cpu().numpy() - g["ssim"]).max() < 2e-3
    # thresholds of inference_video.py:231,246 on the recorded cases: identical -> static, zero frame -> scene cut
    assert got[0] > pkg.scene.STATIC_THRESHOLD and got[-1] < pkg.scene.SCENE_CUT_THRESHOLD
    # 1080p frames (padded to 1088 internally) against the torch restatement
    gen = torch.Generator().manual_seed(3)
    lo = torch.rand(1, 3, 1080 // 8 + 2, 1920 // 8 + 2, generator=gen)
    big = torch.nn.functional.interpolate(lo, size=(1096, 1936), mode="bicubic", align_corners=False).clamp(0, 1)
    a, b = big[:, :, 8:1088, 8:1928].contiguous(), big[:, :, 5:1085, 11:1931].contiguous()
    ref = T.scene_ssim(a, b, 32)
    assert abs(float(pkg.scene_ssim(a.cuda(), b.cuda(), 32)[0]) - ref) < 2e-5
    assert abs(float(pkg.scene_ssim(a.cuda(), a.cuda(), 32)[0]) - 1.0) < 1e-6
    with pytest.raises(RuntimeError):
        pkg.scene_ssim(a, b)                                            # CPU tensors: no fallback
