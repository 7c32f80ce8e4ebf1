"""PyTorch-CPU (fp32) restatement of the reference hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's arithmetic is a sequence of ATen operators; this module issues the same operator sequence
(F.conv2d, F.conv_transpose2d, F.pixel_shuffle, F.interpolate, F.grid_sample, F.leaky_relu, torch.sigmoid) on
a plain state dict, following the reference line by line:
  train_log/RIFE_HDv3.py:56-60, train_log/IFNet_HDv3.py:33-43,53-54,78-90,119-169, model/warplayer.py:8-22.
It is what `bench.py`'s cpu_baseline / `--impl reference` leg times on the host cores (the unmodified reference
lives at /root/reference only in the build container and cannot travel to the GPU box), and what the GPU tests
use as the fp32 reference at sizes where the numpy oracle (oracle/rife_oracle.py) would take too long.  It is
pinned against the same golden fixtures (tests/test_oracle_golden.py).  Never imported by the product.
"""
import math

import torch
import torch.nn.functional as F


def warp(ten_input, ten_flow):
    """model/warplayer.py:8-22 (device = cpu, default dtype float32)."""
    b, _, h, w = ten_flow.shape
    hor = torch.linspace(-1.0, 1.0, w).view(1, 1, 1, w).expand(b, -1, h, -1)
    ver = torch.linspace(-1.0, 1.0, h).view(1, 1, h, 1).expand(b, -1, -1, w)
    grid = torch.cat([hor, ver], 1)
    fl = torch.cat([ten_flow[:, 0:1] / ((ten_input.shape[3] - 1.0) / 2.0),
                    ten_flow[:, 1:2] / ((ten_input.shape[2] - 1.0) / 2.0)], 1)
    g = (grid + fl).permute(0, 2, 3, 1)
    return F.grid_sample(input=ten_input, grid=g, mode="bilinear", padding_mode="border", align_corners=True)


def head(sd, x, p="encode."):
    x = F.leaky_relu(F.conv2d(x, sd[p + "cnn0.weight"], sd[p + "cnn0.bias"], 2, 1), 0.2)
    x = F.leaky_relu(F.conv2d(x, sd[p + "cnn1.weight"], sd[p + "cnn1.bias"], 1, 1), 0.2)
    x = F.leaky_relu(F.conv2d(x, sd[p + "cnn2.weight"], sd[p + "cnn2.bias"], 1, 1), 0.2)
    return F.conv_transpose2d(x, sd[p + "cnn3.weight"], sd[p + "cnn3.bias"], 2, 1)


def ifblock(sd, p, x, flow, scale):
    x = F.interpolate(x, scale_factor=1. / scale, mode="bilinear", align_corners=False)
    if flow is not None:
        flow = F.interpolate(flow, scale_factor=1. / scale, mode="bilinear", align_corners=False) * 1. / scale
        x = torch.cat((x, flow), 1)
    f = x
    for j in range(2):
        f = F.leaky_relu(F.conv2d(f, sd[f"{p}conv0.{j}.0.weight"], sd[f"{p}conv0.{j}.0.bias"], 2, 1), 0.2)
    for j in range(8):
        c = F.conv2d(f, sd[f"{p}convblock.{j}.conv.weight"], sd[f"{p}convblock.{j}.conv.bias"], 1, 1)
        f = F.leaky_relu(c * sd[f"{p}convblock.{j}.beta"] + f, 0.2)
    t = F.pixel_shuffle(F.conv_transpose2d(f, sd[p + "lastconv.0.weight"], sd[p + "lastconv.0.bias"], 2, 1), 2)
    t = F.interpolate(t, scale_factor=scale, mode="bilinear", align_corners=False)
    return t[:, :4] * scale, t[:, 4:5], t[:, 5:]


def ifnet_forward(sd, x, timestep=0.5, scale_list=(16, 8, 4, 2, 1)):
    ch = x.shape[1] // 2
    img0, img1 = x[:, :ch], x[:, ch:]
    if not torch.is_tensor(timestep):
        timestep = (x[:, :1].clone() * 0 + 1) * timestep
    else:
        timestep = timestep.repeat(1, 1, img0.shape[2], img0.shape[3])
    f0, f1 = head(sd, img0[:, :3]), head(sd, img1[:, :3])
    flow_list, merged, mask_list = [], [], []
    w0, w1, flow, mask, feat = img0, img1, None, None, None
    for i in range(5):
        p = f"block{i}."
        if flow is None:
            flow, mask, feat = ifblock(sd, p, torch.cat((img0[:, :3], img1[:, :3], f0, f1, timestep), 1), None, scale_list[i])
        else:
            wf0, wf1 = warp(f0, flow[:, :2]), warp(f1, flow[:, 2:4])
            fd, m0, feat = ifblock(sd, p, torch.cat((w0[:, :3], w1[:, :3], wf0, wf1, timestep, mask, feat), 1), flow, scale_list[i])
            mask = m0
            flow = flow + fd
        mask_list.append(mask)
        flow_list.append(flow)
        w0, w1 = warp(img0, flow[:, :2]), warp(img1, flow[:, 2:4])
        merged.append((w0, w1))
    m = torch.sigmoid(mask)
    merged[4] = w0 * m + w1 * (1 - m)
    return flow_list, mask_list[4], merged


@torch.no_grad()
def model_inference(sd, img0, img1, timestep=0.5, scale=1.0):
    imgs = torch.cat((img0, img1), 1)
    scale_list = [16 / scale, 8 / scale, 4 / scale, 2 / scale, 1 / scale]
    return ifnet_forward(sd, imgs, timestep, scale_list)[2][-1]


# ----------------------------------------------------------------------------- ECCV-paper network (config 5)
def _cp(sd, p, x, stride=1):
    return F.prelu(F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], stride, 1), sd[p + "1.weight"])


def eccv_ifblock(sd, p, x, flow, scale, two_r=False):
    """model/IFNet.py:39-51; two_r: model/IFNet_2R.py:24,48-49 (conv0[0] stride 1, tmp up-sampled by `scale` only)."""
    if scale != 1:
        x = F.interpolate(x, scale_factor=1. / scale, mode="bilinear", align_corners=False)
    if flow is not None:
        flow = F.interpolate(flow, scale_factor=1. / scale, mode="bilinear", align_corners=False) * 1. / scale
        x = torch.cat((x, flow), 1)
    x = _cp(sd, p + "conv0.1.", _cp(sd, p + "conv0.0.", x, 1 if two_r else 2), 2)
    y = x
    for j in range(8):
        y = _cp(sd, f"{p}convblock.{j}.", y)
    x = y + x
    tmp = F.conv_transpose2d(x, sd[p + "lastconv.weight"], sd[p + "lastconv.bias"], 2, 1)
    up = scale if two_r else scale * 2
    tmp = F.interpolate(tmp, scale_factor=up, mode="bilinear", align_corners=False)
    return tmp[:, :4] * up, tmp[:, 4:5]


def _conv2(sd, p, x, stride=2):
    return _cp(sd, p + "conv2.", _cp(sd, p + "conv1.", x, stride))


def eccv_contextnet(sd, x, flow, p="contextnet.", two_r=False):
    """model/refine.py:44-57; two_r: model/refine_2R.py:40,45-48 (conv1 stride 1, f1 warped by the full-resolution flow)."""
    out = []
    for k in range(1, 5):
        x = _conv2(sd, f"{p}conv{k}.", x, 1 if (two_r and k == 1) else 2)
        if not (two_r and k == 1):
            flow = F.interpolate(flow, scale_factor=0.5, mode="bilinear", align_corners=False, recompute_scale_factor=False) * 0.5
        out.append(warp(x, flow))
    return out


def eccv_unet(sd, img0, img1, w0, w1, mask, flow, c0, c1, p="unet.", two_r=False):
    """model/refine.py:72-82; two_r: model/refine_2R.py:63,71 (down0 stride 1, the output conv stride 2)."""
    def deconv(q, x):
        return F.prelu(F.conv_transpose2d(x, sd[q + "0.weight"], sd[q + "0.bias"], 2, 1), sd[q + "1.weight"])
    s0 = _conv2(sd, p + "down0.", torch.cat((img0, img1, w0, w1, mask, flow), 1), 1 if two_r else 2)
    s1 = _conv2(sd, p + "down1.", torch.cat((s0, c0[0], c1[0]), 1))
    s2 = _conv2(sd, p + "down2.", torch.cat((s1, c0[1], c1[1]), 1))
    s3 = _conv2(sd, p + "down3.", torch.cat((s2, c0[2], c1[2]), 1))
    x = deconv(p + "up0.", torch.cat((s3, c0[3], c1[3]), 1))
    x = deconv(p + "up1.", torch.cat((x, s2), 1))
    x = deconv(p + "up2.", torch.cat((x, s1), 1))
    x = deconv(p + "up3.", torch.cat((x, s0), 1))
    return torch.sigmoid(F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], 2 if two_r else 1, 1))


@torch.no_grad()
def eccv_ifnet_forward(sd, x, scale=(4, 2, 1), timestep=None, two_r=False):
    """model/IFNet.py:63-108, inference branch; with `timestep` (float or (B,1,1,1) tensor) the arbitrary-timestep
    variant model/IFNet_m.py:63-112, whose blocks take a constant timestep plane after img1; two_r: the
    model/IFNet_2R.py + model/refine_2R.py variant (same 190 tensors, different strides)."""
    img0, img1 = x[:, :3], x[:, 3:6]
    tp = () if timestep is None else ((x[:, :1].clone() * 0 + 1) * timestep,)       # IFNet_m.py:64
    flow_list, merged, mask_list = [], [], []
    w0, w1, flow, mask = img0, img1, None, None
    for i in range(3):
        p = f"block{i}."
        if flow is not None:
            fd, md = eccv_ifblock(sd, p, torch.cat((img0, img1) + tp + (w0, w1, mask), 1), flow, scale[i], two_r)
            flow, mask = flow + fd, mask + md
        else:
            flow, mask = eccv_ifblock(sd, p, torch.cat((img0, img1) + tp, 1), None, scale[i], two_r)
        mask_list.append(torch.sigmoid(mask))
        flow_list.append(flow)
        w0, w1 = warp(img0, flow[:, :2]), warp(img1, flow[:, 2:4])
        merged.append((w0, w1))
    for i in range(3):
        merged[i] = merged[i][0] * mask_list[i] + merged[i][1] * (1 - mask_list[i])
    c0 = eccv_contextnet(sd, img0, flow[:, :2], two_r=two_r)
    c1 = eccv_contextnet(sd, img1, flow[:, 2:4], two_r=two_r)
    tmp = eccv_unet(sd, img0, img1, w0, w1, mask, flow, c0, c1, two_r=two_r)
    merged[2] = torch.clamp(merged[2] + (tmp[:, :3] * 2 - 1), 0, 1)
    return flow_list, mask_list[2], merged


@torch.no_grad()
def eccv_model_inference(sd, img0, img1, scale=1.0, scale_list=(4, 2, 1), timestep=None):
    sl = [s * 1.0 / scale for s in scale_list]
    return eccv_ifnet_forward(sd, torch.cat((img0, img1), 1), sl, timestep)[2][2]


# ----------------------------------------------------------------------------- scene-change detector (video loop)
def ssim_matlab(img1, img2):
    """model/pytorch_msssim/__init__.py:81-134 for frames in [0, 1] (L = 1): colour images as 3-D volumes, 11^3 Gaussian
    window (sigma 1.5), replicate padding, mean of the SSIM map."""
    g = torch.tensor([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t())
    win = (w2.unsqueeze(2) @ g.t()).expand(1, 1, 11, 11, 11).contiguous()
    a, b = img1.unsqueeze(1), img2.unsqueeze(1)
    f = lambda v: F.conv3d(F.pad(v, (5, 5, 5, 5, 5, 5), mode="replicate"), win)
    mu1, mu2 = f(a), f(b)
    s1, s2, s12 = f(a * a) - mu1 * mu1, f(b * b) - mu2 * mu2, f(a * b) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    v1, v2 = 2.0 * s12 + c2, s1 + s2 + c2
    return (((2 * mu1 * mu2 + c1) * v1) / ((mu1 * mu1 + mu2 * mu2 + c1) * v2)).mean()


@torch.no_grad()
def scene_ssim(i0, i1, pad_multiple=32):
    """inference_video.py:199-203,226-228: zero-pad to the multiple, bilinear-resize to 32x32, ssim_matlab of the first
    three channels.  i0 / i1: (1, C, h, w) float frames."""
    h, w = i0.shape[2:]
    ph, pw = (h + pad_multiple - 1) // pad_multiple * pad_multiple, (w + pad_multiple - 1) // pad_multiple * pad_multiple
    p0, p1 = F.pad(i0, (0, pw - w, 0, ph - h)), F.pad(i1, (0, pw - w, 0, ph - h))
    s0 = F.interpolate(p0, (32, 32), mode="bilinear", align_corners=False)
    s1 = F.interpolate(p1, (32, 32), mode="bilinear", align_corners=False)
    return float(ssim_matlab(s0[:, :3], s1[:, :3]))
