import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import AMP, seeded_model, torch_state_dict
from oracle import torch_ref as T
pkg = importlib.import_module("eccv2022-rife_b200")
rng = np.random.default_rng(22)
lo = torch.from_numpy(rng.random((2, 3, 12, 20), dtype=np.float32))
big = torch.nn.functional.interpolate(lo, size=(90, 150), mode="bicubic", align_corners=False).clamp(0, 1)
f0 = (big * 255).byte().permute(0, 2, 3, 1).contiguous()
f1 = (big.roll(3, 3) * 255).byte().permute(0, 2, 3, 1).contiguous()
m = seeded_model(pkg, AMP, "fp32")
out = m.flownet.run(f0.cuda(), f1.cuda(), 0.5, 1.0, want_all=False).cpu()
a = torch.nn.functional.pad(f0.permute(0, 3, 1, 2).float() / 255., (0, 42, 0, 38))
b = torch.nn.functional.pad(f1.permute(0, 3, 1, 2).float() / 255., (0, 42, 0, 38))
reff = T.model_inference(torch_state_dict(m), a, b)[:, :, :90, :150]
ref = (reff * 255).byte().permute(0, 2, 3, 1)
d = (out.int() - ref.int()).abs()
print("max", d.max().item(), "frac>1", (d > 1).float().mean().item())
idx = (d > 1).nonzero()
print(idx[:10], idx.shape)
# float path with same padded inputs
outf = m.inference(a.cuda(), b.cuda()).cpu()[:, :, :90, :150]
print("float path vs ref", (outf - reff).abs().max().item())
o2 = (outf * 255).byte().permute(0, 2, 3, 1)
print("float->u8 vs ref", (o2.int() - ref.int()).abs().max().item(), "u8 path vs float->u8", (out.int() - o2.int()).abs().max().item())
for k in range(min(5, idx.shape[0])):
    bb, y, x, c = idx[k].tolist(); print(idx[k].tolist(), out[bb, y, x, c].item(), ref[bb, y, x, c].item(), reff[bb, c, y, x].item())
print("---- per-sample")
for i in range(2):
    oi = m.flownet.run(f0[i:i+1].contiguous().cuda(), f1[i:i+1].contiguous().cuda(), 0.5, 1.0, want_all=False).cpu()
    print(i, "single vs ref", (oi[0].int() - ref[i].int()).abs().max().item(), "single vs batched", (oi[0].int() - out[i].int()).abs().max().item())
# which sample does out[1] resemble?
for name, cand in (("ref0", ref[0]), ("f0[1]", f0[1]), ("f1[1]", f1[1]), ("f0[0]", f0[0])):
    print(name, (out[1].int() - cand.int()).abs().float().mean().item())
