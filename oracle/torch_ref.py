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
