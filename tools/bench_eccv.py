"""Times the ECCV-paper network (BASELINE config 5: batch of 448x256 pairs, benchmark/Vimeo90K.py shape) on cuda:0
and the torch-CPU restatement on the host cores."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import torch_ref as T
pkg = importlib.import_module("eccv2022-rife_b200")
torch.manual_seed(1234)
m = pkg.eccv.Model(); m.eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
a, b = torch.rand(B, 3, 256, 448, device="cuda"), torch.rand(B, 3, 256, 448, device="cuda")
for name, (x0, x1) in (("fp32", (a, b)), ("fp16", (a.half(), b.half()))):
    for _ in range(3):
        m.inference(x0, x1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = pkg._lib.load().rife_launch_count()
    e0.record()
    for _ in range(5):
        m.inference(x0, x1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"GPU {name}: batch {B}: {ms:.2f} ms/batch, {B / ms * 1e3:.0f} pairs/s, {253614.7 * 256 * 448 * B / ms / 1e9:.1f} TFLOP/s, launches/batch {(pkg._lib.load().rife_launch_count() - n0) // 5}")
sd = {k: v.detach().cpu().float() for k, v in m.flownet.state_dict().items()}
ac, bc = a[:4].cpu(), b[:4].cpu()
T.eccv_model_inference(sd, ac, bc)
t0 = time.perf_counter(); T.eccv_model_inference(sd, ac, bc); dt = time.perf_counter() - t0
print(f"CPU torch fp32 ({os.cpu_count()} threads): {4 / dt:.1f} pairs/s")
