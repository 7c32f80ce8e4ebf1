"""Shape fuzz on cuda:0: random batch / size / scale / channel count / frame dtype through the fp16 fast path and the fp32
path of the same weights; every result must be finite and the two paths must agree within the reduced-precision bar.
Exercises ragged tiles (sizes that are not multiples of the kernels' tile shapes), uint8 frames with in-kernel pad / crop,
RGBA, multi-timestep and video-order batching."""
import importlib, os, random, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("eccv2022-rife_b200")
torch.manual_seed(1234)
m16 = pkg.Model(precision="fp16"); m16.eval()
m32 = pkg.Model(precision="fp32"); m32.eval()
m32.flownet.load_state_dict(m16.flownet.state_dict())
with torch.no_grad():
    for net in (m16.flownet, m32.flownet):
        for i in range(5):
            getattr(net, f"block{i}").lastconv[0].weight.mul_(4.0)
        net.invalidate()
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
worst = 0.0
for it in range(N):
    scale = rnd.choice([1.0, 1.0, 1.0, 0.5, 0.25, 2.0])
    mult = int(64 / scale)
    B = rnd.choice([1, 1, 2, 3, 5])
    C = rnd.choice([3, 3, 3, 4])
    kind = rnd.choice(["f16", "f16", "u8", "pad", "multi", "seq"])
    H = mult * rnd.randint(1, max(1, 640 // mult)); W = mult * rnd.randint(1, max(1, 1024 // mult))
    g = torch.Generator().manual_seed(it)
    lo = torch.rand(B + 1, C, H // 16 + 3, W // 16 + 3, generator=g)
    big = torch.nn.functional.interpolate(lo, size=(H + 16, W + 16), mode="bicubic", align_corners=False).clamp(0, 1)
    a = big[:B, :, 8:8 + H, 8:8 + W].contiguous().cuda(); b = big[1:, :, 5:5 + H, 11:11 + W].contiguous().cuda()
    t = rnd.choice([0.5, 0.5, 0.25, 0.8])
    if kind == "u8" and C == 3:
        h, w = H - rnd.randint(0, mult - 1), W - 8 * rnd.randint(0, (mult - 1) // 8)
        ua = (a[:, :, :h, :w] * 255).byte().permute(0, 2, 3, 1).contiguous(); ub = (b[:, :, :h, :w] * 255).byte().permute(0, 2, 3, 1).contiguous()
        o16 = m16.flownet.run(ua, ub, t, scale, False, True).float() / 255
        fa, fb = ua.permute(0, 3, 1, 2).float() / 255, ub.permute(0, 3, 1, 2).float() / 255
        o32 = m32.flownet.run(fa.contiguous(), fb.contiguous(), t, scale, False, True).permute(0, 2, 3, 1)
        tol = 2e-2 + 1 / 255
    elif kind == "pad":
        h, w = H - rnd.randint(0, mult - 1), W - rnd.randint(0, mult - 1)
        o16 = m16.inference(a[:, :, :h, :w].half().contiguous(), b[:, :, :h, :w].half().contiguous(), t, scale, pad=True).float()
        o32 = m32.inference(a[:, :, :h, :w].contiguous(), b[:, :, :h, :w].contiguous(), t, scale, pad=True)
        tol = 2e-2
    elif kind == "multi":
        o16 = m16.inference_multi(a.half(), b.half(), (0.25, 0.75), scale).float(); o32 = m32.inference_multi(a, b, (0.25, 0.75), scale)
        tol = 2e-2
    elif kind == "seq" and B >= 2:
        o16 = m16.inference_sequence(a.half(), (t,), scale).float(); o32 = m32.inference_sequence(a, (t,), scale)
        tol = 2e-2
    else:
        o16 = m16.inference(a.half(), b.half(), t, scale).float(); o32 = m32.inference(a, b, t, scale)
        tol = 2e-2
    torch.cuda.synchronize()
    assert o16.shape == o32.shape, (o16.shape, o32.shape)
    assert torch.isfinite(o16).all() and torch.isfinite(o32).all(), (it, kind)
    err = float((o16 - o32).abs().max())
    worst = max(worst, err)
    flag = "" if err < tol else "   <-- OVER"
    print(f"{it:3d} {kind:5s} B={B} C={C} {H}x{W} scale={scale} t={t}: max|fp16-fp32| = {err:.4f}{flag}", flush=True)
    assert err < 2 * tol, "fp16 and fp32 paths disagree"
print("worst", worst)
