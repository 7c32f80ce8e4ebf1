"""1080p fp16 throughput of the caller-side batching modes on cuda:0:
independent pairs / video-order sequence (Head once per frame) / 3 timesteps per pair (4X) / both."""
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
H, W = 1088, 1920
fr = torch.rand(10, 3, H, W, device="cuda").half()
ts3 = (0.25, 0.5, 0.75)


def timeit(name, fn, nout, iters=10):
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
    print(f"{name:58s} {ms:7.3f} ms  {nout / ms * 1e3:8.1f} frames/s")


timeit("2X independent pairs: inference(8 pairs)", lambda: m.inference(fr[:8], fr[1:9]), 8)
timeit("2X video order: inference_sequence(9 frames)", lambda: m.inference_sequence(fr[:9]), 8)
timeit("4X independent: 3 x inference(3 pairs, t)", lambda: [m.inference(fr[:3], fr[1:4], t) for t in ts3], 9)
timeit("4X batched as 9 virtual pairs: inference(9, t tensor)",
       lambda: m.inference(fr[:3].repeat_interleave(3, 0), fr[1:4].repeat_interleave(3, 0),
                           torch.tensor(ts3 * 3).reshape(9, 1, 1, 1)), 9)
timeit("4X inference_multi(3 pairs, 3 timesteps)", lambda: m.inference_multi(fr[:3], fr[1:4], ts3), 9)
timeit("4X video order: inference_sequence(4 frames, 3 timesteps)", lambda: m.inference_sequence(fr[:4], ts3), 9)
