"""Is a small-batch call CPU-bound?  Host time to enqueue N calls vs GPU time to finish them (1080p and 720p, 1 pair)."""
import importlib, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("eccv2022-rife_b200")
torch.manual_seed(1234)
m = pkg.Model(precision="fp16"); m.eval()
for (h, w, P) in ((1088, 1920, 1), (768, 1280, 1), (256, 448, 1), (1088, 1920, 4)):
    a, b = torch.rand(P, 3, h, w, device="cuda").half(), torch.rand(P, 3, h, w, device="cuda").half()
    for _ in range(5):
        m.inference(a, b)
    torch.cuda.synchronize()
    N = 50
    t0 = time.perf_counter()
    for _ in range(N):
        m.inference(a, b)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{h}x{w} P={P}: host enqueue {1e3 * (t1 - t0) / N:.3f} ms/call, to completion {1e3 * (t2 - t0) / N:.3f} ms/call")
