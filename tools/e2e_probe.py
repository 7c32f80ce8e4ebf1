"""Where the end-to-end gap comes from: device-resident fp16 NCHW vs device-resident uint8 HWC 1080-row frames (same
kernels + u8 conversion / pad / crop) vs the pipelined host path (adds the PCIe copies)."""
import importlib, sys, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("eccv2022-rife_b200")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
STEPS = 15
torch.manual_seed(1234)
m = pkg.Model(precision="fp16"); m.eval()
dev = torch.device("cuda", 0)


def timed(fn, n=STEPS, warm=4):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


sets = [(torch.rand(P, 3, 1088, 1920, device=dev).half(), torch.rand(P, 3, 1088, 1920, device=dev).half()) for _ in range(3)]
ms = timed(lambda i: m.inference(*sets[i % 3]))
print(f"P={P} fp16 NCHW 1088 device-resident: {ms:.3f} ms  {P / ms * 1e3:.0f} frames/s")
u8 = [((a[:, :, :1080] * 255).byte().permute(0, 2, 3, 1).contiguous(), (b[:, :, :1080] * 255).byte().permute(0, 2, 3, 1).contiguous()) for a, b in sets]
ms = timed(lambda i: m.flownet.run(u8[i % 3][0], u8[i % 3][1], 0.5, 1.0, False, True))
print(f"P={P} uint8 HWC 1080 device-resident:  {ms:.3f} ms  {P / ms * 1e3:.0f} frames/s")
h8 = [(a.cpu().pin_memory(), b.cpu().pin_memory()) for a, b in u8]
outs = [torch.empty((P, 1080, 1920, 3), dtype=torch.uint8).pin_memory() for _ in range(3)]
streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
cur = torch.cuda.current_stream()


def host_step(i):
    m.inference_host(h8[i % 3][0], h8[i % 3][1], outs[i % 3], slot=i % 3, streams=streams)


for i in range(6):
    host_step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in streams:
    s.wait_stream(cur)
for i in range(STEPS):
    host_step(i)
for s in streams:
    cur.wait_stream(s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / STEPS
print(f"P={P} uint8 HWC 1080 pinned host, pipelined: {ms:.3f} ms  {P / ms * 1e3:.0f} frames/s")

# ---- interference only: device-resident uint8 inference while independent copies of the same size run on other streams
d_in = [torch.empty_like(u8[0][0]) for _ in range(2)]
d_out = torch.empty((P, 1080, 1920, 3), dtype=torch.uint8, device=dev)
for mode in ("h2d", "d2h", "both"):
    def step(i):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(streams[0]):
                d_in[0].copy_(h8[i % 3][0], non_blocking=True)
                d_in[1].copy_(h8[i % 3][1], non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(streams[2]):
                outs[i % 3].copy_(d_out, non_blocking=True)
        m.flownet.run(u8[i % 3][0], u8[i % 3][1], 0.5, 1.0, False, True)
    for s in streams:
        s.wait_stream(cur)
    ms = timed(step)
    torch.cuda.synchronize()
    print(f"P={P} uint8 device-resident + independent {mode} copies of one step's frames: {ms:.3f} ms  {P / ms * 1e3:.0f} frames/s")
