"""CPU oracle for the RIFE HDv3 (v4.25) IFNet inference hot path -- TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module, and only as the checker / the timed CPU baseline.  The product
(`eccv2022-rife_b200`) never imports it and has no CPU fallback.

This is a plain-numpy (float32) restatement of the reference algorithm.  Citations are
`path:line` inside the reference repository (Dk0071942/ECCV2022-RIFE):

  * train_log/RIFE_HDv3.py:56-60    Model.inference      -> `model_inference`
  * train_log/IFNet_HDv3.py:119-169 IFNet.forward        -> `ifnet_forward`
  * train_log/IFNet_HDv3.py:78-90   IFBlock.forward      -> `ifblock_forward`
  * train_log/IFNet_HDv3.py:33-43   Head.forward         -> `head_forward`
  * train_log/IFNet_HDv3.py:53-54   ResConv.forward      -> inside `ifblock_forward`
  * model/warplayer.py:8-22         warp                 -> `warp`

The arithmetic itself lives in PyTorch (an un-vendored dependency of the reference,
`requirements.txt:4` `torch>=1.6.0`; the build container and GPU box have torch 2.11.0+cu128).
The ATen operators used at those call sites are restated here from their published
semantics: conv2d / conv_transpose2d (cross-correlation, NCHW, OIHW / IOHW weights),
leaky_relu, pixel_shuffle, upsample_bilinear2d(align_corners=False, scale_factor given),
grid_sampler_2d(bilinear, border, align_corners=True), linspace, sigmoid.

Pinning: the reference ships no tests or golden vectors.  The pins are
the fixtures under tests/golden/*.npz, produced by tests/golden/make_golden.py by running the
unmodified reference in the build container; tests/test_oracle_golden.py checks this oracle
against every one of them.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------- elementary ops
def leaky_relu(x, slope=0.2):
    """nn.LeakyReLU(0.2) (IFNet_HDv3.py:13,31,51)."""
    return np.where(x >= 0, x, x * F32(slope)).astype(F32)


def sigmoid(x):
    """torch.sigmoid (IFNet_HDv3.py:158)."""
    x = x.astype(F32)
    return (F32(1) / (F32(1) + np.exp(-x))).astype(F32)


def conv2d(x, w, b, stride=1, pad=1):
    """nn.Conv2d forward: x (B,Ci,H,W), w (Co,Ci,kh,kw), b (Co,) -> (B,Co,Ho,Wo)."""
    x = np.ascontiguousarray(x, dtype=F32)
    B, Ci, H, W = x.shape
    Co, Ci2, kh, kw = w.shape
    assert Ci == Ci2
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, (kh, kw), axis=(2, 3))  # B,Ci,H',W',kh,kw
    win = win[:, :, ::stride, ::stride][:, :, :Ho, :Wo]
    cols = np.ascontiguousarray(win.transpose(0, 2, 3, 1, 4, 5)).reshape(B * Ho * Wo, Ci * kh * kw)
    out = cols @ np.ascontiguousarray(w.reshape(Co, -1).T, dtype=F32)
    if b is not None:
        out = out + b.astype(F32)[None, :]
    return np.ascontiguousarray(out.reshape(B, Ho, Wo, Co).transpose(0, 3, 1, 2), dtype=F32)


def conv_transpose2d_k4s2p1(x, w, b):
    """nn.ConvTranspose2d(Ci, Co, 4, 2, 1): w is (Ci,Co,4,4).

    out[co, 2*iy-1+ky, 2*ix-1+kx] += x[ci, iy, ix] * w[ci, co, ky, kx]  (SURVEY 8a row L).
    """
    x = np.ascontiguousarray(x, dtype=F32)
    B, Ci, H, W = x.shape
    Ci2, Co, kh, kw = w.shape
    assert Ci == Ci2 and kh == 4 and kw == 4
    big = np.zeros((B, Co, 2 * H + 2, 2 * W + 2), dtype=F32)  # index = 2*i + k  (offset +1 vs output)
    xm = x.transpose(0, 2, 3, 1).reshape(-1, Ci)
    for ky in range(4):
        for kx in range(4):
            contrib = (xm @ w[:, :, ky, kx].astype(F32)).reshape(B, H, W, Co).transpose(0, 3, 1, 2)
            big[:, :, ky:ky + 2 * H:2, kx:kx + 2 * W:2] += contrib
    out = big[:, :, 1:1 + 2 * H, 1:1 + 2 * W]
    if b is not None:
        out = out + b.astype(F32)[None, :, None, None]
    return np.ascontiguousarray(out, dtype=F32)


def pixel_shuffle2(x):
    """nn.PixelShuffle(2): ps[k, 2y+dy, 2x+dx] = x[4k + 2dy + dx, y, x] (IFNet_HDv3.py:75)."""
    B, C, H, W = x.shape
    k = C // 4
    y = x.reshape(B, k, 2, 2, H, W).transpose(0, 1, 4, 2, 5, 3).reshape(B, k, 2 * H, 2 * W)
    return np.ascontiguousarray(y, dtype=F32)


def _src_index(out_size, in_size, scale_factor):
    """ATen area_pixel_compute_source_index, align_corners=False, scale_factor provided."""
    ratio = F32(1.0 / scale_factor)  # compute_scales_value: 1/scale when a scale is given
    d = np.arange(out_size, dtype=F32)
    src = ratio * (d + F32(0.5)) - F32(0.5)
    src = np.maximum(src, F32(0)).astype(F32)
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = np.where(i0 < in_size - 1, i0 + 1, i0)
    l1 = (src - i0.astype(F32)).astype(F32)
    l0 = (F32(1) - l1).astype(F32)
    return i0, i1, l0, l1


def interpolate_bilinear(x, scale_factor):
    """F.interpolate(x, scale_factor=s, mode='bilinear', align_corners=False) (IFNet_HDv3.py:79,81,86)."""
    x = np.ascontiguousarray(x, dtype=F32)
    B, C, H, W = x.shape
    Ho = int(np.floor(H * float(scale_factor)))
    Wo = int(np.floor(W * float(scale_factor)))
    y0, y1, hl0, hl1 = _src_index(Ho, H, scale_factor)
    x0, x1, wl0, wl1 = _src_index(Wo, W, scale_factor)
    top = x[:, :, y0, :]
    bot = x[:, :, y1, :]
    wl0 = wl0[None, None, None, :]
    wl1 = wl1[None, None, None, :]
    t = wl0 * top[:, :, :, x0] + wl1 * top[:, :, :, x1]
    b = wl0 * bot[:, :, :, x0] + wl1 * bot[:, :, :, x1]
    out = hl0[None, None, :, None] * t + hl1[None, None, :, None] * b
    return out.astype(F32)


def linspace_m1_1(n):
    """torch.linspace(-1.0, 1.0, n) in float32 (symmetric two-sided formula of ATen RangeFactories)."""
    if n == 1:
        return np.array([-1.0], dtype=F32)
    start, end = F32(-1.0), F32(1.0)
    step = F32((end - start) / F32(n - 1))
    idx = np.arange(n, dtype=np.int64)
    lo = (start + step * idx.astype(F32)).astype(F32)
    hi = (end - step * (n - 1 - idx).astype(F32)).astype(F32)
    return np.where(idx < n // 2, lo, hi).astype(F32)


def grid_sample_border_ac(x, grid):
    """F.grid_sample(bilinear, padding_mode='border', align_corners=True); grid (B,H,W,2) in [-1,1]."""
    x = np.ascontiguousarray(x, dtype=F32)
    B, C, H, W = x.shape
    gx = grid[..., 0].astype(F32)
    gy = grid[..., 1].astype(F32)
    ix = ((gx + F32(1)) / F32(2)) * F32(W - 1)
    iy = ((gy + F32(1)) / F32(2)) * F32(H - 1)
    ix = np.minimum(F32(W - 1), np.maximum(ix, F32(0))).astype(F32)
    iy = np.minimum(F32(H - 1), np.maximum(iy, F32(0))).astype(F32)
    x0 = np.floor(ix).astype(F32)
    y0 = np.floor(iy).astype(F32)
    x1 = x0 + F32(1)
    y1 = y0 + F32(1)
    nw = (x1 - ix) * (y1 - iy)
    ne = (ix - x0) * (y1 - iy)
    sw = (x1 - ix) * (iy - y0)
    se = (ix - x0) * (iy - y0)
    x0i, y0i, x1i, y1i = (a.astype(np.int64) for a in (x0, y0, x1, y1))

    def tap(yi, xi, wgt):
        ok = (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
        yc = np.clip(yi, 0, H - 1)
        xc = np.clip(xi, 0, W - 1)
        bi = np.arange(B)[:, None, None]
        v = x[bi, :, yc, xc]  # (B,H,W,C)
        return v * (wgt * ok.astype(F32))[..., None]

    out = tap(y0i, x0i, nw) + tap(y0i, x1i, ne) + tap(y1i, x0i, sw) + tap(y1i, x1i, se)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2), dtype=F32)


def warp(ten_input, ten_flow, grids=None):
    """model/warplayer.py:8-22: backward warp with the reference's normalised-coordinate order.

    `grids=(gx[W], gy[H])` optionally supplies the base grids.  ATen's *vectorised* CPU linspace differs from
    the scalar formula of `linspace_m1_1` by 1 ulp at ~40% of the entries (and depends on the host's SIMD
    width), which moves warped noise images by up to ~7e-6; pass torch.linspace's values to remove that."""
    ten_input = np.asarray(ten_input, dtype=F32)
    ten_flow = np.asarray(ten_flow, dtype=F32)
    B, _, H, W = ten_flow.shape
    gx, gy = (linspace_m1_1(W), linspace_m1_1(H)) if grids is None else (np.asarray(grids[0], F32), np.asarray(grids[1], F32))
    hor = np.broadcast_to(gx[None, None, :], (B, H, W))                     # :11-12
    ver = np.broadcast_to(gy[None, :, None], (B, H, W))                     # :13-14
    fx = ten_flow[:, 0] / F32((ten_input.shape[3] - 1.0) / 2.0)            # :18
    fy = ten_flow[:, 1] / F32((ten_input.shape[2] - 1.0) / 2.0)            # :19
    g = np.stack([(hor + fx).astype(F32), (ver + fy).astype(F32)], axis=-1)  # :21
    return grid_sample_border_ac(ten_input, g)                               # :22


# ----------------------------------------------------------------------------- network
def _p(sd, key):
    return np.asarray(sd[key], dtype=F32)


def head_forward(sd, x, prefix="encode."):
    """Head.forward (IFNet_HDv3.py:33-43): 3 convs + LeakyReLU, ConvTranspose2d(16,4,4,2,1)."""
    x = leaky_relu(conv2d(x, _p(sd, prefix + "cnn0.weight"), _p(sd, prefix + "cnn0.bias"), stride=2))
    x = leaky_relu(conv2d(x, _p(sd, prefix + "cnn1.weight"), _p(sd, prefix + "cnn1.bias")))
    x = leaky_relu(conv2d(x, _p(sd, prefix + "cnn2.weight"), _p(sd, prefix + "cnn2.bias")))
    return conv_transpose2d_k4s2p1(x, _p(sd, prefix + "cnn3.weight"), _p(sd, prefix + "cnn3.bias"))


def ifblock_forward(sd, prefix, x, flow, scale):
    """IFBlock.forward (IFNet_HDv3.py:78-90)."""
    x = interpolate_bilinear(x, 1.0 / scale)                                             # :79
    if flow is not None:
        flow = (interpolate_bilinear(flow, 1.0 / scale) * F32(1.0) / F32(scale)).astype(F32)  # :81
        x = np.concatenate([x, flow], axis=1)                                            # :82
    f = x
    for j in range(2):                                                                    # :83 conv0
        f = leaky_relu(conv2d(f, _p(sd, f"{prefix}conv0.{j}.0.weight"), _p(sd, f"{prefix}conv0.{j}.0.bias"), stride=2))
    for j in range(8):                                                                    # :84 ResConv x8
        c = conv2d(f, _p(sd, f"{prefix}convblock.{j}.conv.weight"), _p(sd, f"{prefix}convblock.{j}.conv.bias"))
        f = leaky_relu(c * _p(sd, f"{prefix}convblock.{j}.beta") + f)                    # :54
    t = conv_transpose2d_k4s2p1(f, _p(sd, prefix + "lastconv.0.weight"), _p(sd, prefix + "lastconv.0.bias"))
    t = pixel_shuffle2(t)                                                                 # :85
    t = interpolate_bilinear(t, scale)                                                    # :86
    return (t[:, :4] * F32(scale)).astype(F32), t[:, 4:5], t[:, 5:]                     # :87-89


def ifnet_forward(sd, x, timestep=0.5, scale_list=(16, 8, 4, 2, 1)):
    """IFNet.forward (IFNet_HDv3.py:119-169), inference branch (training=False, fastmode=True)."""
    x = np.asarray(x, dtype=F32)
    ch = x.shape[1] // 2
    img0, img1 = x[:, :ch], x[:, ch:]                                                     # :121-123
    B, _, H, W = img0.shape
    if np.ndim(timestep) == 0:
        tplane = np.full((B, 1, H, W), F32(timestep), dtype=F32)                          # :125
    else:
        tplane = np.broadcast_to(np.asarray(timestep, dtype=F32).reshape(B, 1, 1, 1), (B, 1, H, W)).astype(F32)  # :127
    f0 = head_forward(sd, img0[:, :3])                                                    # :128
    f1 = head_forward(sd, img1[:, :3])                                                    # :129
    flow_list, merged, mask_list = [], [], []
    w0, w1 = img0, img1
    flow = mask = feat = None
    for i in range(5):                                                                    # :139
        pre = f"block{i}."
        if flow is None:
            xin = np.concatenate([img0[:, :3], img1[:, :3], f0, f1, tplane], axis=1)      # :141
            flow, mask, feat = ifblock_forward(sd, pre, xin, None, scale_list[i])
        else:
            wf0 = warp(f0, flow[:, :2])                                                   # :145
            wf1 = warp(f1, flow[:, 2:4])                                                  # :146
            xin = np.concatenate([w0[:, :3], w1[:, :3], wf0, wf1, tplane, mask, feat], axis=1)  # :147
            fd, m0, feat = ifblock_forward(sd, pre, xin, flow, scale_list[i])
            mask = m0                                                                     # :151
            flow = (flow + fd).astype(F32)                                                # :152
        mask_list.append(mask)
        flow_list.append(flow)
        w0 = warp(img0, flow[:, :2])                                                      # :155
        w1 = warp(img1, flow[:, 2:4])                                                     # :156
        merged.append((w0, w1))
    m = sigmoid(mask)                                                                     # :158
    merged[4] = (w0 * m + w1 * (F32(1) - m)).astype(F32)                                  # :159
    return flow_list, mask_list[4], merged


def model_inference(sd, img0, img1, timestep=0.5, scale=1.0):
    """Model.inference (RIFE_HDv3.py:56-60)."""
    imgs = np.concatenate([np.asarray(img0, F32), np.asarray(img1, F32)], axis=1)
    scale_list = [16 / scale, 8 / scale, 4 / scale, 2 / scale, 1 / scale]
    _, _, merged = ifnet_forward(sd, imgs, timestep, scale_list)
    return merged[-1]


def state_dict_to_numpy(sd):
    """torch state_dict (optionally 'module.'-prefixed, RIFE_HDv3.py:37-45) -> {key: float32 ndarray}."""
    out = {}
    for k, v in sd.items():
        k = k.replace("module.", "")
        out[k] = v.detach().cpu().numpy().astype(F32) if hasattr(v, "detach") else np.asarray(v, F32)
    return out


# ----------------------------------------------------------------------------- ECCV-paper network (config 5)
def prelu(x, slope):
    """nn.PReLU(C): per-channel slope (model/IFNet.py:17, model/refine.py:15,21)."""
    s = np.asarray(slope, F32).reshape(1, -1, 1, 1)
    return np.where(x >= 0, x, x * s).astype(F32)


def _conv_prelu(sd, p, x, stride=1):
    return prelu(conv2d(x, _p(sd, p + "0.weight"), _p(sd, p + "0.bias"), stride, 1), _p(sd, p + "1.weight"))


def eccv_ifblock(sd, p, x, flow, scale):
    """IFBlock.forward of model/IFNet.py:39-51."""
    if scale != 1:
        x = interpolate_bilinear(x, 1.0 / scale)                                         # :41
    if flow is not None:
        flow = (interpolate_bilinear(flow, 1.0 / scale) * F32(1.0) / F32(scale)).astype(F32)  # :43
        x = np.concatenate([x, flow], 1)
    x = _conv_prelu(sd, p + "conv0.0.", x, 2)                                             # :45
    x = _conv_prelu(sd, p + "conv0.1.", x, 2)
    y = x
    for j in range(8):
        y = _conv_prelu(sd, f"{p}convblock.{j}.", y)
    x = (y + x).astype(F32)                                                               # :46
    tmp = conv_transpose2d_k4s2p1(x, _p(sd, p + "lastconv.weight"), _p(sd, p + "lastconv.bias"))  # :47
    tmp = interpolate_bilinear(tmp, scale * 2)                                            # :48
    return (tmp[:, :4] * F32(scale) * F32(2)).astype(F32), tmp[:, 4:5]                   # :49-50


def _conv2(sd, p, x):
    """Conv2 of model/refine.py:24-33."""
    return _conv_prelu(sd, p + "conv2.", _conv_prelu(sd, p + "conv1.", x, 2))


def eccv_contextnet(sd, x, flow, p="contextnet."):
    """model/refine.py:44-57."""
    out = []
    for k in range(1, 5):
        x = _conv2(sd, f"{p}conv{k}.", x)
        flow = (interpolate_bilinear(flow, 0.5) * F32(0.5)).astype(F32)
        out.append(warp(x, flow))
    return out


def eccv_unet(sd, img0, img1, w0, w1, mask, flow, c0, c1, p="unet."):
    """model/refine.py:72-82."""
    def deconv(q, x):
        return prelu(conv_transpose2d_k4s2p1(x, _p(sd, q + "0.weight"), _p(sd, q + "0.bias")), _p(sd, q + "1.weight"))
    s0 = _conv2(sd, p + "down0.", np.concatenate([img0, img1, w0, w1, mask, flow], 1))
    s1 = _conv2(sd, p + "down1.", np.concatenate([s0, c0[0], c1[0]], 1))
    s2 = _conv2(sd, p + "down2.", np.concatenate([s1, c0[1], c1[1]], 1))
    s3 = _conv2(sd, p + "down3.", np.concatenate([s2, c0[2], c1[2]], 1))
    x = deconv(p + "up0.", np.concatenate([s3, c0[3], c1[3]], 1))
    x = deconv(p + "up1.", np.concatenate([x, s2], 1))
    x = deconv(p + "up2.", np.concatenate([x, s1], 1))
    x = deconv(p + "up3.", np.concatenate([x, s0], 1))
    return sigmoid(conv2d(x, _p(sd, p + "conv.weight"), _p(sd, p + "conv.bias"), 1, 1))


def eccv_ifnet_forward(sd, x, scale=(4, 2, 1)):
    """IFNet.forward of model/IFNet.py:63-108, inference branch (x = cat(img0, img1), no ground truth)."""
    x = np.asarray(x, F32)
    img0, img1 = x[:, :3], x[:, 3:6]
    flow_list, merged, mask_list = [], [], []
    w0, w1, flow, mask = img0, img1, None, None
    for i in range(3):
        p = f"block{i}."
        if flow is not None:
            fd, md = eccv_ifblock(sd, p, np.concatenate([img0, img1, w0, w1, mask], 1), flow, scale[i])  # :76
            flow = (flow + fd).astype(F32)
            mask = (mask + md).astype(F32)
        else:
            flow, mask = eccv_ifblock(sd, p, np.concatenate([img0, img1], 1), None, scale[i])           # :80
        mask_list.append(sigmoid(mask))
        flow_list.append(flow)
        w0, w1 = warp(img0, flow[:, :2]), warp(img1, flow[:, 2:4])                                       # :83-84
        merged.append((w0, w1))
    for i in range(3):
        merged[i] = (merged[i][0] * mask_list[i] + merged[i][1] * (F32(1) - mask_list[i])).astype(F32)   # :99
    c0 = eccv_contextnet(sd, img0, flow[:, :2])                                                          # :103
    c1 = eccv_contextnet(sd, img1, flow[:, 2:4])
    tmp = eccv_unet(sd, img0, img1, w0, w1, mask, flow, c0, c1)                                          # :105
    res = (tmp[:, :3] * F32(2) - F32(1)).astype(F32)
    merged[2] = np.clip(merged[2] + res, 0, 1).astype(F32)                                               # :107
    return flow_list, mask_list[2], merged


def eccv_model_inference(sd, img0, img1, scale=1.0, scale_list=(4, 2, 1)):
    """Model.inference of model/RIFE.py:56-65 (TTA=False)."""
    sl = [s * 1.0 / scale for s in scale_list]
    return eccv_ifnet_forward(sd, np.concatenate([np.asarray(img0, F32), np.asarray(img1, F32)], 1), sl)[2][2]
