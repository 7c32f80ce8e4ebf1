"""Print the fp16-mode error margins of the ECCV-paper path against the fp32 goldens (run on the GPU box)."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from conftest import load_golden, seeded_eccv  # noqa: E402

pkg = importlib.import_module("eccv2022-rife_b200")
g = load_golden("eccv_cases")
m = seeded_eccv(pkg)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
psnr = lambda a, b: float(-10 * np.log10(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2) + 1e-20))
x = torch.cat((cu(g["a_img0"]), cu(g["a_img1"])), 1).half()
flows, mask, merged, *_ = m.flownet(x, [4, 2, 1])
for i in range(3):
    o = merged[i].float().cpu().numpy()
    print("a merged", i, np.abs(o - g[f"a_merged{i}"]).max(), psnr(o, g[f"a_merged{i}"]),
          "flow", np.abs(flows[i].float().cpu().numpy() - g[f"a_flow{i}"]).max(), np.abs(g[f"a_flow{i}"]).max())
print("a mask", np.abs(mask.float().cpu().numpy() - g["a_mask"]).max())
o = m.inference(cu(g["b_img0"]).half(), cu(g["b_img1"]).half(), scale=0.5, scale_list=[4, 2, 1]).float().cpu().numpy()
print("b", np.abs(o - g["b_out"]).max(), psnr(o, g["b_out"]))
o = m.inference(cu(g["c_img0"]), cu(g["c_img1"]), scale_list=[4, 2, 1]).float().cpu().numpy()
r = g["c_out"].astype(np.float32)
print("c", np.abs(o - r).max(), psnr(o, r))
