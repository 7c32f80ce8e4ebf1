"""The reference's own device path as a yardstick ("existing GPU kernel" bar): the same ATen operator sequence
(oracle/torch_ref.py) issued through eager PyTorch / cuDNN on cuda:0, cudnn.benchmark on as in inference_video.py:76-79,
warp grid cached per size as model/warplayer.py:9-16 does.  1080p (1088x1920), fp32 and the reference's fp16 mode."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import torch_ref as T
import torch.nn.functional as F

torch.backends.cudnn.enabled = True
torch.backends.cudnn.benchmark = True
_grid = {}


def warp_cuda(ten_input, ten_flow):
    k = (str(ten_flow.device), str(ten_flow.size()), ten_flow.dtype)
    if k not in _grid:
        b, _, h, w = ten_flow.shape
        hor = torch.linspace(-1.0, 1.0, w, device=ten_flow.device).view(1, 1, 1, w).expand(b, -1, h, -1)
        ver = torch.linspace(-1.0, 1.0, h, device=ten_flow.device).view(1, 1, h, 1).expand(b, -1, -1, w)
        _grid[k] = torch.cat([hor, ver], 1).to(ten_flow.dtype)
    fl = torch.cat([ten_flow[:, 0:1] / ((ten_input.shape[3] - 1.0) / 2.0),
                    ten_flow[:, 1:2] / ((ten_input.shape[2] - 1.0) / 2.0)], 1)
    g = (_grid[k] + fl).permute(0, 2, 3, 1)
    return F.grid_sample(input=ten_input, grid=g, mode="bilinear", padding_mode="border", align_corners=True)


T.warp = warp_cuda
pkg = importlib.import_module("eccv2022-rife_b200")
torch.manual_seed(1234)
m = pkg.Model(precision="fp16"); m.eval()
sd32 = {k: v.detach().float().cuda() for k, v in m.flownet.state_dict().items()}
sd16 = {k: v.half() for k, v in sd32.items()}
for P in (1, 8):
    a, b = torch.rand(P, 3, 1088, 1920, device="cuda"), torch.rand(P, 3, 1088, 1920, device="cuda")
    for name, sd, x0, x1 in (("fp32", sd32, a, b), ("fp16", sd16, a.half(), b.half())):
        for _ in range(3):
            T.model_inference(sd, x0, x1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            T.model_inference(sd, x0, x1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"eager PyTorch / cuDNN on B200, {name}, {P} pair(s) of 1088x1920: {ms:.2f} ms  {P / ms * 1e3:.0f} frames/s", flush=True)
    ms_ours = None
    for _ in range(3):
        m.inference(a.half(), b.half())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.inference(a.half(), b.half())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"this implementation, fp16 mode, {P} pair(s): {ms:.2f} ms  {P / ms * 1e3:.0f} frames/s", flush=True)
