"""Device-resident throughput of BASELINE.json's other configurations on cuda:0 (they are parity-test cases; bench.py measures
the 1080p 2X headline): config 2 (720p 2X, fp16, B in {1, 4, 16}), config 3 (1080p 4X: 3 timesteps per pair), config 4
(4K 2X at scale = 0.5), plus single-pair 1080p latency.  FLOP / byte figures per pair from SURVEY.md section 8d."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("eccv2022-rife_b200")
torch.manual_seed(1234)
m = pkg.model.Model(precision="fp16")
m.eval()


def timeit(name, fn, nout, flop_per_out, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:64s} {ms:8.3f} ms  {nout / ms * 1e3:8.1f} frames/s  {flop_per_out * nout / ms / 1e9:6.1f} TFLOP/s")


def flop(h, w, scale=1.0):
    return h * w * (6064 + 38372.3 * scale * scale)


def frames(n, h, w):
    return torch.rand(n, 3, h, w, device="cuda").half(), torch.rand(n, 3, h, w, device="cuda").half()


for B in (1, 4, 16, 32):
    a, b = frames(B, 768, 1280)
    timeit(f"config 2: 720p (768x1280) 2X, B = {B}", lambda: m.inference(a, b), B, flop(768, 1280))
    del a, b
for B in (1, 4):
    a, b = frames(B, 1088, 1920)
    timeit(f"1080p (1088x1920) 2X, B = {B}", lambda: m.inference(a, b), B, flop(1088, 1920))
a, b = frames(4, 1088, 1920)
timeit("config 3: 1080p 4X, inference_multi(4 pairs, t = .25 .5 .75)", lambda: m.inference_multi(a, b, (0.25, 0.5, 0.75)), 12,
       flop(1088, 1920))
del a, b
for B in (1, 4):
    a, b = frames(B, 2176, 3840)
    timeit(f"config 4: 4K (2176x3840) 2X at scale = 0.5, B = {B}", lambda: m.inference(a, b, 0.5, 0.5), B, flop(2176, 3840, 0.5))
    del a, b
