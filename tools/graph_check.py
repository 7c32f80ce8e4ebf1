"""CUDA-graph replay vs plain launches on the same inputs (run twice: default env and RIFE_GRAPH_MAX=0, compare files)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import AMP, seeded_model
pkg = importlib.import_module("eccv2022-rife_b200")
m = seeded_model(pkg, AMP, "fp16")
g = torch.Generator().manual_seed(21)
xs = [torch.rand(1, 3, 128, 192, generator=g).cuda().half() for _ in range(7)]
outs = [m.inference(xs[i], xs[i + 1], 0.25 + 0.1 * i).cpu() for i in range(6)]
same_t = [m.inference(xs[i], xs[i + 1], 0.5).cpu() for i in range(6)]
torch.save((outs, same_t), sys.argv[1])
if len(sys.argv) > 2:
    ref_o, ref_s = torch.load(sys.argv[2])
    for i in range(6):
        print(i, "varying t: max diff", float((outs[i].float() - ref_o[i].float()).abs().max()),
              " fixed t:", float((same_t[i].float() - ref_s[i].float()).abs().max()))
