"""Prints fp16- / bf16-mode error margins (max-abs / PSNR vs the fp32 goldens and the torch-CPU restatement).
    python tools/fp16_margins.py [fp16|bf16]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import AMP, load_golden, seeded_model, torch_state_dict, psnr
from oracle import torch_ref as T
pkg = importlib.import_module("eccv2022-rife_b200")
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
g = load_golden("inference_variants")
PREC = sys.argv[1] if len(sys.argv) > 1 else "fp16"
print("precision mode:", PREC)
m, ma = seeded_model(pkg, 1.0, PREC), seeded_model(pkg, AMP, PREC)
def rep(name, out, ref): print(f"{name:28s} max-abs {np.abs(out-ref).max():.5f}  PSNR {psnr(out, ref):.1f} dB")
rep("a noise amp t=.3", ma.inference(cu(g["a_img0"]), cu(g["a_img1"]), 0.3).cpu().numpy(), g["a_out"])
rep("b smooth amp scale .5", ma.inference(cu(g["b_img0"]), cu(g["b_img1"]), 0.5, 0.5).cpu().numpy(), g["b_out"])
rep("d noise amp tensor-t", ma.inference(cu(g["d_img0"]), cu(g["d_img1"]), cu(g["d_t"])).cpu().numpy(), g["d_out"])
rep("e smooth default", m.inference(cu(g["e_img0"]), cu(g["e_img1"])).cpu().numpy(), g["e_out"])
d = load_golden("demo_256x448")
i0 = cu(d["img0_u8"].transpose(2, 0, 1)[None]).float() / 255.; i1 = cu(d["img1_u8"].transpose(2, 0, 1)[None]).float() / 255.
rep("demo default", m.inference(i0, i1).cpu().numpy(), d["out"])
rep("demo amp", ma.inference(i0, i1).cpu().numpy(), d["out_amp"])
gen = torch.Generator().manual_seed(10)
lo = torch.rand(1, 3, 768 // 8 + 2, 1280 // 8 + 2, generator=gen)
big = torch.nn.functional.interpolate(lo, size=(768 + 16, 1280 + 16), mode="bicubic", align_corners=False).clamp(0, 1)
a, b = big[:, :, 8:8 + 768, 8:8 + 1280].contiguous(), big[:, :, 5:5 + 768, 11:11 + 1280].contiguous()
ref = T.model_inference(torch_state_dict(ma), a, b).numpy()
rep("720p smooth amp half-io", ma.inference(a.cuda().half(), b.cuda().half()).float().cpu().numpy(), ref)
