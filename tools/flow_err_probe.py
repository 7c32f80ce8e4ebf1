import sys, importlib, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import AMP, seeded_model, load_golden
pkg = importlib.import_module('eccv2022-rife_b200')
cu = lambda a: torch.from_numpy(a).cuda()
g = load_golden("ifnet_full_64x128_amp")
net = seeded_model(pkg, AMP, "fp32").flownet
x = torch.cat((cu(g["img0"]), cu(g["img1"])), 1)
flows, mask, merged = net(x, 0.5, [16, 8, 4, 2, 1])
for i in range(5):
    r = g[f"flow{i}"]; e = np.abs(flows[i].cpu().numpy() - r).max()
    print("ifnet_full flow", i, "err", e, "absmax", np.abs(r).max(), "rel", e / max(1, np.abs(r).max()))
g = load_golden("scale_up_cases")
m32 = seeded_model(pkg, AMP, "fp32")
a, b = cu(g["s2_img0"]), cu(g["s2_img1"])
flows, mask, merged = m32.flownet(torch.cat((a, b), 1), 0.3, [8, 4, 2, 1, 0.5])
for k, t in (("s2_flow4", flows[4]), ("s2_flow3", flows[3]), ("s2_mask", mask)):
    r = g[k]; e = np.abs(t.cpu().numpy() - r).max()
    print(k, "err", e, "absmax", np.abs(r).max(), "rel", e / max(1, np.abs(r).max()))
# ECCV-paper nets (tests/test_gpu_eccv.py)
from conftest import seeded_eccv, seeded_eccv_m, seeded_eccv_2r
for name, mk, key in (("eccv_cases", seeded_eccv, "a"), ("eccv_2r_cases", seeded_eccv_2r, "a")):
    g = load_golden(name)
    m = mk(pkg)
    net = m.flownet if hasattr(m, "flownet") else m
    x = torch.cat((cu(g["a_img0"]), cu(g["a_img1"])), 1)
    flows = net(x, [4, 2, 1])[0]
    for i in range(3):
        r = g[f"a_flow{i}"]; e = np.abs(flows[i].cpu().numpy() - r).max()
        print(name, "flow", i, "err", e, "absmax", np.abs(r).max(), "rel", e / max(1, np.abs(r).max()))
