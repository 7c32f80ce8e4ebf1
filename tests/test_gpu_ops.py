"""Op-level parity of the CUDA kernels (through the C ABI) against the numpy oracle and the golden vectors."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rife_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _stream():
    return torch.cuda.current_stream().cuda_stream


def test_warp_golden(pkg):
    g = load_golden("warp_cases")
    for name in ("small", "large", "huge"):
        x = torch.from_numpy(g[f"{name}_x"]).to(DEV)
        fl = torch.from_numpy(g[f"{name}_flow"]).to(DEV)
        y = pkg.warp(x, fl).cpu().numpy()
        # bit-level: same base grid (torch.linspace on the host), same arithmetic order as model/warplayer.py
        assert np.abs(y - g[f"{name}_y"]).max() < 1e-5, name


def test_warp_fp16_and_errors(pkg):
    g = load_golden("warp_cases")
    x = torch.from_numpy(g["small_x"]).to(DEV).half()
    fl = torch.from_numpy(g["small_flow"]).to(DEV).half()
    y = pkg.warp(x, fl).float().cpu().numpy()
    ref = O.warp(x.float().cpu().numpy(), fl.float().cpu().numpy())
    assert np.abs(y - ref).max() < 2e-3
    with pytest.raises(RuntimeError):
        pkg.warp(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))          # CPU tensors: no fallback
    with pytest.raises(ValueError):
        pkg.warp(torch.zeros(1, 3, 8, 8, device=DEV), torch.zeros(1, 2, 8, 4, device=DEV))


@pytest.mark.parametrize("num,den", [(1, 2), (1, 4), (1, 16), (1, 1), (2, 1), (8, 1), (16, 1)])
def test_resize_bilinear(pkg, num, den):
    lib = pkg._lib.load()
    rng = np.random.default_rng(num * 100 + den)
    x = rng.standard_normal((2, 5, 32, 64)).astype(np.float32)
    xd = torch.from_numpy(x).to(DEV)
    out = torch.empty((2, 5, 32 * num // den, 64 * num // den), device=DEV)
    pkg._lib.check(lib.rife_resize_bilinear(xd.data_ptr(), out.data_ptr(), 2, 5, 32, 64, num, den, 3.0, _stream()))
    ref = O.interpolate_bilinear(x, num / den) * np.float32(3.0)
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-6 * max(1.0, float(np.abs(ref).max()))   # ~1 ulp (FMA contraction)


def _conv(pkg, x, w, b, stride, act, beta=None, res=None, precision=0):
    lib = pkg._lib.load()
    B, Ci, H, W = x.shape
    Co = w.shape[0]
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    xd, wd, bd, bed, rd = dev(x), dev(w), dev(b), dev(beta), dev(res)
    y = torch.empty((B, Co, H // stride, W // stride), device=DEV)
    ptr = lambda t: None if t is None else t.data_ptr()
    pkg._lib.check(lib.rife_conv3x3(ptr(xd), ptr(wd), ptr(bd), ptr(bed), ptr(rd), y.data_ptr(), B, Ci, H, W, Co, stride,
                                    act, precision, _stream()))
    torch.cuda.synchronize()
    return y.cpu().numpy()


@pytest.mark.parametrize("ci,co,h,w,stride", [(16, 16, 24, 40, 1), (28, 16, 32, 48, 2), (15, 96, 16, 16, 2),
                                              (32, 32, 17, 30, 1), (96, 96, 20, 36, 1), (192, 192, 9, 15, 1),
                                              (3, 16, 32, 32, 2), (64, 128, 16, 32, 2)])
def test_conv3x3_fp32(pkg, ci, co, h, w, stride):
    rng = np.random.default_rng(ci * 7 + co)
    x = rng.standard_normal((2, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, 3, 3)) / np.sqrt(9 * ci)).astype(np.float32)
    b = rng.standard_normal(co).astype(np.float32)
    y = _conv(pkg, x, wt, b, stride, 1)
    ref = O.leaky_relu(O.conv2d(x, wt, b, stride, 1))
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() < 2e-5


def test_resconv_fp32(pkg):
    """ResConv.forward (IFNet_HDv3.py:53-54): lrelu(conv(x) * beta + x)."""
    rng = np.random.default_rng(5)
    c = 64
    x = rng.standard_normal((1, c, 20, 28)).astype(np.float32)
    wt = (rng.standard_normal((c, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32)
    b = rng.standard_normal(c).astype(np.float32)
    beta = rng.standard_normal(c).astype(np.float32)
    y = _conv(pkg, x, wt, b, 1, 1, beta=beta, res=x)
    ref = O.leaky_relu(O.conv2d(x, wt, b, 1, 1) * beta[None, :, None, None] + x)
    assert np.abs(y - ref).max() < 2e-5


@pytest.mark.parametrize("ci,co,ps", [(32, 52, 1), (16, 4, 0), (192, 52, 1)])
def test_conv_transpose_fp32(pkg, ci, co, ps):
    """lastconv = ConvTranspose2d(c, 52, 4, 2, 1) + PixelShuffle(2) (IFNet_HDv3.py:73-76); Head.cnn3 (:30)."""
    lib = pkg._lib.load()
    rng = np.random.default_rng(ci + co)
    x = rng.standard_normal((2, ci, 10, 14)).astype(np.float32)
    wt = (rng.standard_normal((ci, co, 4, 4)) / np.sqrt(4 * ci)).astype(np.float32)
    b = rng.standard_normal(co).astype(np.float32)
    xd, wd, bd = (torch.from_numpy(a).to(DEV) for a in (x, wt, b))
    shape = (2, co // 4, 40, 56) if ps else (2, co, 20, 28)
    y = torch.empty(shape, device=DEV)
    pkg._lib.check(lib.rife_conv_transpose4x4s2(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), 2, ci, 10, 14,
                                                co, ps, 0, _stream()))
    ref = O.conv_transpose2d_k4s2p1(x, wt, b)
    if ps:
        ref = O.pixel_shuffle2(ref)
    assert np.abs(y.cpu().numpy() - ref).max() < 2e-5


# ---------------------------------------------------------------------------------------------- tcgen05 (fp16) convs
def _h(a):
    return a.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("ci,co,h,w,stride", [(16, 16, 24, 40, 1), (64, 64, 16, 32, 1), (32, 32, 17, 30, 1),
                                              (96, 96, 20, 36, 1), (128, 128, 12, 20, 1), (192, 192, 9, 15, 1),
                                              (28, 16, 32, 48, 2), (15, 96, 16, 16, 2), (3, 16, 32, 32, 2),
                                              (64, 128, 16, 32, 2), (32, 64, 8, 8, 2), (16, 32, 64, 96, 2)])
@pytest.mark.parametrize("prec", [1, 3])   # 1 = pipeline policy (halo kernel when the layer fits), 3 = NHWC kernel only
def test_conv3x3_fp16_tcgen05(pkg, ci, co, h, w, stride, prec):
    rng = np.random.default_rng(ci * 7 + co)
    x = _h(rng.standard_normal((2, ci, h, w)).astype(np.float32))
    wt = _h((rng.standard_normal((co, ci, 3, 3)) / np.sqrt(9 * ci)).astype(np.float32))
    b = rng.standard_normal(co).astype(np.float32)
    y = _conv(pkg, x, wt, b, stride, 1, precision=prec)
    ref = O.leaky_relu(O.conv2d(x, wt, b, stride, 1))
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() < 4e-3 * max(1.0, float(np.abs(ref).max()))   # fp16 output rounding (2^-11 relative)


@pytest.mark.parametrize("prec,c,hh", [(1, 64, 20), (3, 64, 20), (1, 32, 20), (1, 96, 20), (1, 96, 40), (3, 96, 20),
                                       (1, 128, 20), (1, 128, 36), (1, 192, 20)])
def test_resconv_fp16_tcgen05(pkg, prec, c, hh):
    rng = np.random.default_rng(5)
    x = _h(rng.standard_normal((2, c, hh, 28)).astype(np.float32))
    wt = _h((rng.standard_normal((c, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32))
    b = rng.standard_normal(c).astype(np.float32)
    beta = rng.standard_normal(c).astype(np.float32)
    y = _conv(pkg, x, wt, b, 1, 1, beta=beta, res=x, precision=prec)
    ref = O.leaky_relu(O.conv2d(x, wt, b, 1, 1) * beta[None, :, None, None] + x)
    assert np.abs(y - ref).max() < 4e-3 * max(1.0, float(np.abs(ref).max()))


@pytest.mark.parametrize("prec", [1, 3])
@pytest.mark.parametrize("ci,co,ps", [(32, 52, 1), (16, 4, 0), (192, 52, 1), (64, 52, 1), (96, 52, 1), (128, 52, 1)])
def test_conv_transpose_fp16_tcgen05(pkg, ci, co, ps, prec):
    lib = pkg._lib.load()
    rng = np.random.default_rng(ci + co)
    x = _h(rng.standard_normal((2, ci, 10, 14)).astype(np.float32))
    wt = _h((rng.standard_normal((ci, co, 4, 4)) / np.sqrt(4 * ci)).astype(np.float32))
    b = rng.standard_normal(co).astype(np.float32)
    xd, wd, bd = (torch.from_numpy(a).to(DEV) for a in (x, wt, b))
    shape = (2, co // 4, 40, 56) if ps else (2, co, 20, 28)
    y = torch.empty(shape, device=DEV)
    pkg._lib.check(lib.rife_conv_transpose4x4s2(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), 2, ci, 10, 14,
                                                co, ps, prec, _stream()))
    ref = O.conv_transpose2d_k4s2p1(x, wt, b)
    if ps:
        ref = O.pixel_shuffle2(ref)
    assert np.abs(y.cpu().numpy() - ref).max() < 4e-3 * max(1.0, float(np.abs(ref).max()))


# ---------------------------------------------------------------------------------------------- bf16 operands (RIFE_PRECISION_BF16)
def _b(a):
    """round to bf16 (nearest even) and back"""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return t.to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("ci,co,h,w,stride", [(16, 16, 24, 40, 1), (32, 32, 17, 30, 1), (64, 64, 16, 32, 1), (96, 96, 20, 36, 1),
                                              (192, 192, 9, 15, 1), (64, 128, 16, 32, 2), (16, 32, 64, 96, 2)])
@pytest.mark.parametrize("prec", [2, 4])   # 2 = bf16, pipeline policy (halo kernel when the layer fits); 4 = bf16, NHWC kernel only
def test_conv3x3_bf16_tcgen05(pkg, ci, co, h, w, stride, prec):
    """bf16 x bf16 tcgen05 MMAs with fp32 accumulation; the activation output is stored as bf16 (2^-9 relative rounding)."""
    rng = np.random.default_rng(ci * 7 + co)
    x = _b(rng.standard_normal((2, ci, h, w)).astype(np.float32))
    wt = _b((rng.standard_normal((co, ci, 3, 3)) / np.sqrt(9 * ci)).astype(np.float32))
    b = rng.standard_normal(co).astype(np.float32)
    y = _conv(pkg, x, wt, b, stride, 1, precision=prec)
    ref = O.leaky_relu(O.conv2d(x, wt, b, stride, 1))
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() < 8e-3 * max(1.0, float(np.abs(ref).max()))


@pytest.mark.parametrize("prec,c,hh", [(2, 64, 20), (4, 64, 20), (2, 32, 20), (2, 96, 40), (2, 128, 20), (2, 192, 20)])
def test_resconv_bf16_tcgen05(pkg, prec, c, hh):
    rng = np.random.default_rng(5)
    x = _b(rng.standard_normal((2, c, hh, 28)).astype(np.float32))
    wt = _b((rng.standard_normal((c, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32))
    b = rng.standard_normal(c).astype(np.float32)
    beta = rng.standard_normal(c).astype(np.float32)
    y = _conv(pkg, x, wt, b, 1, 1, beta=beta, res=x, precision=prec)
    ref = O.leaky_relu(O.conv2d(x, wt, b, 1, 1) * beta[None, :, None, None] + x)
    assert np.abs(y - ref).max() < 8e-3 * max(1.0, float(np.abs(ref).max()))


def test_conv_bf16_range_beyond_fp16(pkg):
    """What the mode is for: activations far above the fp16 maximum (65504) come through finite and accurate."""
    rng = np.random.default_rng(8)
    x = _b((rng.standard_normal((1, 32, 16, 24)) * 3e4).astype(np.float32))
    wt = _b((rng.standard_normal((32, 32, 3, 3)) / np.sqrt(9 * 32) * 8).astype(np.float32))
    b = rng.standard_normal(32).astype(np.float32)
    y = _conv(pkg, x, wt, b, 1, 1, precision=2)
    ref = O.leaky_relu(O.conv2d(x, wt, b, 1, 1))
    assert float(np.abs(ref).max()) > 2e5 and np.isfinite(y).all()
    assert np.abs(y - ref).max() < 8e-3 * float(np.abs(ref).max())


@pytest.mark.parametrize("ci", [32, 64, 128])
def test_conv_transpose_bf16_tcgen05(pkg, ci):
    """lastconv in bf16 mode: bf16 operands, fp16 tmp pixels out."""
    lib = pkg._lib.load()
    rng = np.random.default_rng(ci)
    x = _b(rng.standard_normal((2, ci, 10, 14)).astype(np.float32))
    wt = _b((rng.standard_normal((ci, 52, 4, 4)) / np.sqrt(4 * ci)).astype(np.float32))
    b = rng.standard_normal(52).astype(np.float32)
    xd, wd, bd = (torch.from_numpy(a).to(DEV) for a in (x, wt, b))
    y = torch.empty((2, 13, 40, 56), device=DEV)
    pkg._lib.check(lib.rife_conv_transpose4x4s2(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), 2, ci, 10, 14,
                                                52, 1, 2, _stream()))
    ref = O.pixel_shuffle2(O.conv_transpose2d_k4s2p1(x, wt, b))
    assert np.abs(y.cpu().numpy() - ref).max() < 4e-3 * max(1.0, float(np.abs(ref).max()))


def test_base_grid_vs_cuda_linspace(pkg):
    """model/warplayer.py:11-14 builds its base grid on the module-level device -- on a GPU box with ATen's CUDA linspace.
    This package pins the CPU kernel's values (the goldens are CPU recordings); the two differ by at most one ulp, and a
    warp with either grid by < 1e-4 on white noise with 20 px flows."""
    from importlib import import_module
    wl = import_module("eccv2022-rife_b200.warplayer")
    rng = np.random.default_rng(3)
    for h, w in ((256, 448), (1088, 1920)):
        gx, gy, dx, dy = wl.base_grids(h, w, torch.device(DEV))
        cx, cy = torch.linspace(-1.0, 1.0, w, device=DEV), torch.linspace(-1.0, 1.0, h, device=DEV)
        assert float((dx - cx).abs().max()) <= 1.2e-7 and float((dy - cy).abs().max()) <= 1.2e-7
    x = torch.from_numpy(rng.random((1, 3, 256, 448), dtype=np.float32)).to(DEV)
    fl = torch.from_numpy((rng.random((1, 2, 256, 448), dtype=np.float32) * 2 - 1) * 20).to(DEV)
    ours = wl.warp(x, fl)
    gxx = torch.linspace(-1.0, 1.0, 448, device=DEV).view(1, 1, 1, 448).expand(1, -1, 256, -1)
    gyy = torch.linspace(-1.0, 1.0, 256, device=DEV).view(1, 1, 256, 1).expand(1, -1, -1, 448)
    grid = torch.cat((gxx, gyy), 1)
    flow_n = torch.cat((fl[:, 0:1] / ((448 - 1.0) / 2.0), fl[:, 1:2] / ((256 - 1.0) / 2.0)), 1)
    ref = torch.nn.functional.grid_sample(x, (grid + flow_n).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border",
                                          align_corners=True)       # the reference's warp, evaluated by ATen on the GPU
    assert float((ours - ref).abs().max()) < 1e-4       # measured 2.5e-5: ATen's CUDA grid_sample vs the reference's CPU arithmetic
