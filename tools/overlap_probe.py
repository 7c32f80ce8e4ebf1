"""Do the latency-bound coarse blocks (0-2) of one sub-batch hide under the bandwidth-bound stages of another?
Two independent models (own weights / plans / workspaces) each run P/2 pairs per call on their own stream, the second
stream offset by a fraction of a step; compare with one model running P pairs per call on one stream.
    python tools/overlap_probe.py [P] [steps]"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("eccv2022-rife_b200")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
H, W = 1088, 1920
dev = torch.device("cuda", 0)


def mk():
    torch.manual_seed(1234)
    m = pkg.Model(precision="fp16")
    m.eval()
    return m


def frames(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, H, W, generator=g).half().to(dev), torch.rand(n, 3, H, W, generator=g).half().to(dev)


def bench_single(p):
    m = mk()
    a, b = frames(p, 1)
    for _ in range(3):
        m.inference(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        m.inference(a, b)
    e1.record()
    torch.cuda.synchronize()
    return p * STEPS / (e0.elapsed_time(e1) / 1e3)


def bench_multi(p_each, nstreams, offsets):
    ms = [mk() for _ in range(nstreams)]
    data = [frames(p_each, 10 + i) for i in range(nstreams)]
    small = frames(max(1, p_each // 4), 99)
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    for m, (a, b) in zip(ms, data):
        for _ in range(2):
            m.inference(a, b)
        m.inference(*small)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    for s in streams:
        s.wait_stream(cur)
    # offsets: stream i first runs offsets[i] small calls so that the streams are out of phase
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            for _ in range(offsets[i]):
                ms[i].inference(*small)
    for _ in range(STEPS):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                ms[i].inference(*data[i])
    for s in streams:
        cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    extra = sum(offsets) * small[0].shape[0]
    return (p_each * nstreams * STEPS + extra) / (e0.elapsed_time(e1) / 1e3)


print(f"single stream, {P} pairs per call:      {bench_single(P):8.1f} frames/s")
print(f"single stream, {P // 2} pairs per call:      {bench_single(P // 2):8.1f} frames/s")
for off in ((0, 0), (0, 1), (0, 2)):
    print(f"two streams x {P // 2} pairs, offsets {off}:   {bench_multi(P // 2, 2, off):8.1f} frames/s")
print(f"three streams x {P // 2} pairs, offsets (0,1,2): {bench_multi(P // 2, 3, (0, 1, 2)):8.1f} frames/s")
print(f"two streams x {P} pairs, offsets (0,2):  {bench_multi(P, 2, (0, 2)):8.1f} frames/s")
