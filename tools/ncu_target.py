"""Workload for ncu captures: warm up, then run ONE inference between cudaProfilerStart/Stop so that
`ncu --profile-from-start off` sees exactly the launches of one pass.

  python tools/ncu_target.py hdv3 [pairs]     1080p 2X (1088x1920), fp16, default 8 pairs
  python tools/ncu_target.py eccv [batch]     ECCV-paper net, 256x448, fp16, default batch 64
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("eccv2022-rife_b200")
mode = sys.argv[1] if len(sys.argv) > 1 else "hdv3"
torch.manual_seed(1234)
if mode == "hdv3":
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    m = pkg.model.Model(precision="fp16")
    m.eval()
    a, b = torch.rand(P, 3, 1088, 1920, device="cuda").half(), torch.rand(P, 3, 1088, 1920, device="cuda").half()
    run = lambda: m.inference(a, b)
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    m = pkg.eccv.Model(precision="fp16")
    m.eval()
    a, b = torch.rand(B, 3, 256, 448, device="cuda").half(), torch.rand(B, 3, 256, 448, device="cuda").half()
    run = lambda: m.inference(a, b)
for _ in range(3):
    run()
torch.cuda.synchronize()
n0 = pkg._lib.load().rife_launch_count()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in the profiled pass:", pkg._lib.load().rife_launch_count() - n0)
