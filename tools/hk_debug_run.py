import importlib, sys, torch
sys.path.insert(0, '.')
pkg = importlib.import_module("eccv2022-rife_b200")
torch.manual_seed(1234)
m = pkg.Model(precision="fp16"); m.eval()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
a = torch.rand(P, 3, 1088, 1920, device="cuda").half(); b = torch.rand(P, 3, 1088, 1920, device="cuda").half()
for _ in range(2):
    m.inference(a, b); torch.cuda.synchronize()
    print("----", file=sys.stderr)
