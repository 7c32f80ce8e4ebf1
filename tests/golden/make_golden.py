#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED reference on CPU (fp32).

Run in the build container only (the reference lives at /root/reference, read-only,
and does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference has no tests / golden vectors of its own, so these
fixtures -- outputs of the reference's own `train_log/RIFE_HDv3.Model.inference`,
`train_log/IFNet_HDv3.IFNet.forward` and `model/warplayer.warp` on seeded inputs -- are
what pins the oracle (`oracle/`) and, through it, the CUDA path.

Weights: `train_log/flownet.pkl` is absent from the container (.MISSING_LARGE_BLOBS), so
weights are the reference's default init under `torch.manual_seed(1234)`.  They are NOT
stored (22.7 MB); instead `weights_seed1234.json` stores per-tensor float64 sums and a
sha256 so that a test can prove that the host-side mirror module regenerates them
bit-exactly from the same seed.  `amp` cases multiply every `block{i}.lastconv.0.weight`
by AMP to push |flow| to tens of pixels and exercise the border clamp of warp().
"""
import hashlib
import json
import os
import sys

import numpy as np

REF = os.environ.get("RIFE_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count() or 1)

from train_log.RIFE_HDv3 import Model  # noqa: E402  (the reference)
from model import warplayer as ref_warplayer  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
AMP = 4.0


def build_model(seed=1234, amp=1.0):
    torch.manual_seed(seed)
    m = Model()
    m.eval()
    if amp != 1.0:
        for i in range(5):
            getattr(m.flownet, f"block{i}").lastconv[0].weight.mul_(amp)
    return m


def rand_frames(seed, b, c, h, w, smooth=False):
    g = torch.Generator().manual_seed(seed)
    if not smooth:
        return torch.rand(b, c, h, w, generator=g), torch.rand(b, c, h, w, generator=g)
    # smooth frames: low-res noise upsampled bicubically, second frame shifted by a few px
    lo = torch.rand(b, c, h // 8 + 2, w // 8 + 2, generator=g)
    big = torch.nn.functional.interpolate(lo, size=(h + 16, w + 16), mode="bicubic", align_corners=False).clamp(0, 1)
    return big[:, :, 8:8 + h, 8:8 + w].contiguous(), big[:, :, 5:5 + h, 11:11 + w].contiguous()


def f32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


def weights_manifest():
    m = build_model(1234)
    sd = m.flownet.state_dict()
    h = hashlib.sha256()
    sums = {}
    for k, v in sd.items():
        a = v.detach().cpu().numpy().astype(np.float32)
        h.update(k.encode())
        h.update(a.tobytes())
        sums[k] = {"shape": list(a.shape), "sum": float(a.astype(np.float64).sum())}
    man = {"seed": 1234, "torch": torch.__version__, "sha256": h.hexdigest(), "n_tensors": len(sd),
           "n_params": int(sum(v.numel() for v in sd.values())), "tensors": sums}
    with open(os.path.join(OUT, "weights_seed1234.json"), "w") as f:
        json.dump(man, f, indent=1)
    print("weights manifest:", man["sha256"], man["n_tensors"], man["n_params"])


def case_demo():
    """Config 1: inference_img.py on demo/I0_0.png + I0_1.png, exp=1 (inference_img.py:57-60, 99)."""
    import cv2
    i0 = cv2.imread(os.path.join(REF, "demo/I0_0.png"), cv2.IMREAD_UNCHANGED)
    i1 = cv2.imread(os.path.join(REF, "demo/I0_1.png"), cv2.IMREAD_UNCHANGED)
    img0 = (torch.tensor(i0.transpose(2, 0, 1)) / 255.).unsqueeze(0)
    img1 = (torch.tensor(i1.transpose(2, 0, 1)) / 255.).unsqueeze(0)
    m = build_model(1234)
    out = m.inference(img0, img1)
    print("demo mean", float(out.mean()), out[0, :, 100, 200].tolist())
    flow_list, mask, merged = m.flownet(torch.cat((img0, img1), 1), 0.5, [16, 8, 4, 2, 1])
    # the amplified-flow variant on the same smooth natural images
    ma = build_model(1234, AMP)
    out_amp = ma.inference(img0, img1)
    fl_amp, _, _ = ma.flownet(torch.cat((img0, img1), 1), 0.5, [16, 8, 4, 2, 1])
    print("demo amp |flow|max", float(fl_amp[-1].abs().max()))
    save("demo_256x448", img0_u8=i0, img1_u8=i1, out=f32(out), flow=f32(flow_list[-1]).astype(np.float16),
         mask=f32(mask).astype(np.float16), out_amp=f32(out_amp), flow_amp=f32(fl_amp[-1]).astype(np.float16))


def case_full_structure():
    """IFNet.forward full return structure on a small random input (IFNet_HDv3.py:119-169)."""
    img0, img1 = rand_frames(7, 1, 3, 64, 128)
    m = build_model(1234, AMP)
    flow_list, mask, merged = m.flownet(torch.cat((img0, img1), 1), 0.5, [16, 8, 4, 2, 1])
    arrs = {"img0": f32(img0), "img1": f32(img1), "mask": f32(mask), "merged4": f32(merged[4])}
    for i in range(5):
        arrs[f"flow{i}"] = f32(flow_list[i])
    for i in range(4):
        arrs[f"merged{i}_0"] = f32(merged[i][0])
        arrs[f"merged{i}_1"] = f32(merged[i][1])
    save("ifnet_full_64x128_amp", **arrs)


def case_inference_variants():
    arrs = {}
    # (a) batch 2, arbitrary timestep, amplified flow, random noise (worst case for warp parity)
    img0, img1 = rand_frames(11, 2, 3, 128, 192)
    arrs["a_img0"], arrs["a_img1"] = f32(img0), f32(img1)
    arrs["a_out"] = f32(build_model(1234, AMP).inference(img0, img1, timestep=0.3))
    # (b) scale=0.5 (H, W multiples of 128), smooth frames
    img0, img1 = rand_frames(12, 1, 3, 128, 256, smooth=True)
    arrs["b_img0"], arrs["b_img1"] = f32(img0), f32(img1)
    arrs["b_out"] = f32(build_model(1234, AMP).inference(img0, img1, 0.5, 0.5))
    # (c) RGBA input (4 channels: IFNet_HDv3.py:121-123 splits C/2, Head sees [:, :3])
    img0, img1 = rand_frames(13, 1, 4, 64, 64)
    arrs["c_img0"], arrs["c_img1"] = f32(img0), f32(img1)
    arrs["c_out"] = f32(build_model(1234).inference(img0, img1))
    # (d) tensor timestep of shape (B,1,1,1) (IFNet_HDv3.py:126-127)
    img0, img1 = rand_frames(14, 2, 3, 64, 64)
    ts = torch.tensor([0.25, 0.75]).view(2, 1, 1, 1)
    arrs["d_img0"], arrs["d_img1"], arrs["d_t"] = f32(img0), f32(img1), f32(ts)
    arrs["d_out"] = f32(build_model(1234, AMP).inference(img0, img1, ts))
    # (e) default-init, smooth, 128x128
    img0, img1 = rand_frames(15, 1, 3, 128, 128, smooth=True)
    arrs["e_img0"], arrs["e_img1"] = f32(img0), f32(img1)
    arrs["e_out"] = f32(build_model(1234).inference(img0, img1))
    save("inference_variants", **arrs)


def case_scale_up():
    """--scale 2 / 4 (inference_video.py:92): scale_list = [8,4,2,1,0.5] / [4,2,1,0.5,0.25] -- the last block(s) up-sample
    their input and work above full resolution (IFNet_HDv3.py:79-81,86-87 with scale < 1)."""
    arrs = {}
    img0, img1 = rand_frames(61, 2, 3, 64, 96, smooth=True)                 # multiples of 64/2 = 32
    arrs["s2_img0"], arrs["s2_img1"] = f32(img0), f32(img1)
    arrs["s2_out"] = f32(build_model(1234, AMP).inference(img0, img1, 0.5, 2.0))
    fl, mk, mg = build_model(1234, AMP).flownet(torch.cat((img0, img1), 1), 0.3, [8, 4, 2, 1, 0.5])
    arrs["s2_flow4"], arrs["s2_flow3"], arrs["s2_mask"], arrs["s2_merged_t03"] = f32(fl[4]), f32(fl[3]), f32(mk), f32(mg[4])
    img0, img1 = rand_frames(62, 1, 3, 48, 80)                              # multiples of 64/4 = 16, white noise
    arrs["s4_img0"], arrs["s4_img1"] = f32(img0), f32(img1)
    arrs["s4_out"] = f32(build_model(1234, AMP).inference(img0, img1, 0.5, 4.0))
    save("scale_up_cases", **arrs)


def case_scale_list():
    """IFNet.forward with scale lists that are NOT the halving pyramid of RIFE_HDv3.py:58 (IFNet_HDv3.py:119,139-148 takes
    any five numbers): repeated levels, a coarse-to-fine-to-coarse order, and a block above full resolution."""
    arrs = {}
    m = build_model(1234, AMP)
    img0, img1 = rand_frames(81, 1, 3, 64, 96, smooth=True)
    arrs["img0"], arrs["img1"] = f32(img0), f32(img1)
    for name, sl in (("a", [8, 8, 4, 2, 1]), ("b", [4, 4, 2, 2, 1]), ("c", [2, 4, 8, 2, 1]), ("d", [8, 4, 2, 2, 0.5])):
        fl, mk, mg = m.flownet(torch.cat((img0, img1), 1), 0.4, sl)
        arrs[name + "_list"] = np.asarray(sl, np.float32)
        arrs[name + "_flow4"], arrs[name + "_flow1"] = f32(fl[4]), f32(fl[1])
        arrs[name + "_mask"], arrs[name + "_merged"] = f32(mk), f32(mg[4])
    save("scale_list_cases", **arrs)


def case_warp():
    """model/warplayer.py:8-22 on random data, small and large flows, C=3 and C=4."""
    arrs = {}
    g = torch.Generator().manual_seed(21)
    for name, (b, c, h, w, mag) in {"small": (2, 3, 48, 80, 2.0), "large": (1, 4, 64, 96, 40.0),
                                    "huge": (1, 3, 32, 32, 200.0)}.items():
        x = torch.rand(b, c, h, w, generator=g)
        fl = (torch.rand(b, 2, h, w, generator=g) * 2 - 1) * mag
        ref_warplayer.backwarp_tenGrid.clear()
        y = ref_warplayer.warp(x, fl)
        arrs[f"{name}_x"], arrs[f"{name}_flow"], arrs[f"{name}_y"] = f32(x), f32(fl), f32(y)
    save("warp_cases", **arrs)


def case_ops():
    """Intermediate ops of IFBlock.forward / Head.forward on one small input, for op-level pins."""
    m = build_model(1234)
    net = m.flownet
    g = torch.Generator().manual_seed(31)
    arrs = {}
    x = torch.rand(1, 3, 64, 64, generator=g)
    arrs["head_x"] = f32(x)
    arrs["head_y"] = f32(net.encode(x))
    xb = torch.rand(1, 24, 64, 64, generator=g)
    fl = (torch.rand(1, 4, 64, 64, generator=g) * 2 - 1) * 3
    f, mk, ft = net.block3(xb, fl, scale=2)
    arrs["blk3_x"], arrs["blk3_flow"] = f32(xb), f32(fl)
    arrs["blk3_oflow"], arrs["blk3_omask"], arrs["blk3_ofeat"] = f32(f), f32(mk), f32(ft)
    x0 = torch.rand(1, 15, 64, 64, generator=g)
    f, mk, ft = net.block0(x0, None, scale=16)
    arrs["blk0_x"] = f32(x0)
    arrs["blk0_oflow"], arrs["blk0_omask"], arrs["blk0_ofeat"] = f32(f), f32(mk), f32(ft)
    save("ops_cases", **arrs)


def case_eccv():
    """BASELINE config 5: the ECCV-paper net (model/RIFE.py Model.inference, model/IFNet.py IFNet.forward)."""
    from model.RIFE import Model as EccvModel
    torch.manual_seed(1234)
    m = EccvModel()
    m.eval()
    sd = m.flownet.state_dict()
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().astype(np.float32).tobytes())
    with open(os.path.join(OUT, "weights_eccv_seed1234.json"), "w") as f:
        json.dump({"seed": 1234, "sha256": h.hexdigest(), "n_tensors": len(sd),
                   "n_params": int(sum(v.numel() for v in sd.values()))}, f, indent=1)
    # amplify lastconv so that |flow| is a few pixels with random weights
    for b in (m.flownet.block0, m.flownet.block1, m.flownet.block2):
        b.lastconv.weight.mul_(AMP)
    arrs = {}
    img0, img1 = rand_frames(41, 2, 3, 64, 96, smooth=True)
    flow_list, mask, merged, _, _, _ = m.flownet(torch.cat((img0, img1), 1), [4, 2, 1])
    arrs["a_img0"], arrs["a_img1"] = f32(img0), f32(img1)
    for i in range(3):
        arrs[f"a_flow{i}"], arrs[f"a_merged{i}"] = f32(flow_list[i]), f32(merged[i])
    arrs["a_mask"] = f32(mask)
    img0, img1 = rand_frames(42, 1, 3, 128, 128)
    arrs["b_img0"], arrs["b_img1"] = f32(img0), f32(img1)
    arrs["b_out"] = f32(m.inference(img0, img1, scale=0.5, scale_list=[4, 2, 1]))
    img0, img1 = rand_frames(43, 1, 3, 256, 448, smooth=True)          # Vimeo90K frame shape (benchmark/Vimeo90K.py)
    arrs["c_img0"], arrs["c_img1"] = f32(img0).astype(np.float16), f32(img1).astype(np.float16)
    arrs["c_out"] = f32(m.inference(img0.half().float(), img1.half().float(), scale_list=[4, 2, 1])).astype(np.float16)
    print("eccv |flow|max", float(flow_list[-1].abs().max()))
    save("eccv_cases", **arrs)


def case_eccv_m():
    """model/RIFE.py Model(arbitrary=True) = model/IFNet_m.py (benchmark/HD_multi_4X.py:16): timestep plane as an
    extra block input."""
    from model.RIFE import Model as EccvModel
    torch.manual_seed(1234)
    m = EccvModel(arbitrary=True)
    m.eval()
    sd = m.flownet.state_dict()
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().astype(np.float32).tobytes())
    with open(os.path.join(OUT, "weights_eccv_m_seed1234.json"), "w") as f:
        json.dump({"seed": 1234, "sha256": h.hexdigest(), "n_tensors": len(sd),
                   "n_params": int(sum(v.numel() for v in sd.values()))}, f, indent=1)
    for b in (m.flownet.block0, m.flownet.block1, m.flownet.block2):
        b.lastconv.weight.mul_(AMP)
    arrs = {}
    img0, img1 = rand_frames(51, 2, 3, 64, 96, smooth=True)
    arrs["img0"], arrs["img1"] = f32(img0), f32(img1)
    for t in (0.25, 0.5, 0.8):
        flow_list, mask, merged, _, _, _ = m.flownet(torch.cat((img0, img1), 1), [4, 2, 1], timestep=t)
        arrs[f"flow2_t{t}"], arrs[f"mask_t{t}"], arrs[f"out_t{t}"] = f32(flow_list[2]), f32(mask), f32(merged[2])
    arrs["out_scale05_t0.25"] = f32(m.inference(img0, img1, scale=0.5, scale_list=[4, 2, 1], timestep=0.25))
    arrs["flow_only_t0.5"] = f32(m.flownet(torch.cat((img0, img1), 1), [4, 2, 1], timestep=0.5, returnflow=True))
    save("eccv_m_cases", **arrs)


def case_eccv_2r():
    """model/IFNet_2R.py + model/refine_2R.py: the stride variant of the ECCV net (IFBlock.conv0[0], Contextnet.conv1 and
    Unet.down0 at stride 1, the Unet's output conv at stride 2).  No class of the reference wires it into a Model, so the
    fixture drives IFNet.forward directly.  Same 190 tensors / default init as model/IFNet.py."""
    from model.IFNet_2R import IFNet as IFNet2R
    torch.manual_seed(1234)
    net = IFNet2R()
    net.eval()
    for b in (net.block0, net.block1, net.block2):
        b.lastconv.weight.mul_(AMP)
    arrs = {}
    img0, img1 = rand_frames(61, 2, 3, 64, 96, smooth=True)
    flow_list, mask, merged, _, _, _ = net(torch.cat((img0, img1), 1), [4, 2, 1])
    arrs["a_img0"], arrs["a_img1"] = f32(img0), f32(img1)
    for i in range(3):
        arrs[f"a_flow{i}"] = f32(flow_list[i])
    arrs["a_merged0"], arrs["a_merged2"], arrs["a_mask"] = f32(merged[0]), f32(merged[2]), f32(mask)
    img0, img1 = rand_frames(62, 1, 3, 128, 128)
    arrs["b_img0"], arrs["b_img1"] = f32(img0), f32(img1)
    arrs["b_out"] = f32(net(torch.cat((img0, img1), 1), [8, 4, 2])[2][2])          # Model.inference(scale=0.5)'s scale list
    print("eccv 2R |flow|max", float(flow_list[-1].abs().max()))
    save("eccv_2r_cases", **arrs)


def case_scene_ssim():
    """The video loop's scene-change statistic (inference_video.py:199-203,226-228): pad to 32, F.interpolate to
    32x32, model/pytorch_msssim ssim_matlab -- on uint8 frames of a few kinds, recorded from the reference's code."""
    from model.pytorch_msssim import ssim_matlab
    rng = np.random.default_rng(77)
    h, w = 135, 250                                           # not multiples of 32: exercises the padding
    base = F.interpolate(torch.from_numpy(rng.random((1, 3, 12, 20), dtype=np.float32)), (h + 20, w + 20), mode="bicubic",
                         align_corners=False).clamp(0, 1)
    fr = lambda dy, dx: (base[0, :, 10 + dy:10 + dy + h, 10 + dx:10 + dx + w] * 255).byte().permute(1, 2, 0).contiguous().numpy()
    noise = lambda: rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    pairs = [(fr(0, 0), fr(0, 0)), (fr(0, 0), fr(0, 1)), (fr(0, 0), fr(3, -5)), (fr(0, 0), fr(10, 10)), (fr(0, 0), noise()),
             (noise(), noise()), (np.zeros((h, w, 3), np.uint8), fr(0, 0))]
    vals = []
    for a, b in pairs:
        ta = torch.from_numpy(a.transpose(2, 0, 1)).unsqueeze(0).float() / 255.
        tb = torch.from_numpy(b.transpose(2, 0, 1)).unsqueeze(0).float() / 255.
        ph, pw = ((h - 1) // 32 + 1) * 32, ((w - 1) // 32 + 1) * 32
        ta, tb = F.pad(ta, (0, pw - w, 0, ph - h)), F.pad(tb, (0, pw - w, 0, ph - h))          # inference_video.py:199-203
        sa = F.interpolate(ta, (32, 32), mode="bilinear", align_corners=False)
        sb = F.interpolate(tb, (32, 32), mode="bilinear", align_corners=False)
        vals.append(float(ssim_matlab(sa[:, :3], sb[:, :3])))
    print("scene ssim", vals)
    save("scene_ssim_cases", img0=np.stack([p[0] for p in pairs]), img1=np.stack([p[1] for p in pairs]),
         ssim=np.asarray(vals, np.float64))


RANGE_K = 2.0e5    # "range" stress: activations of every conv chain scaled by K (fp16 overflows, the result barely moves)
BIGFLOW_AMP = 20.0  # "bigflow" stress: |flow| of 50-200 px through the whole net


def apply_range_stress(net, k=RANGE_K, amp=AMP):
    """LeakyReLU is positively homogeneous: multiplying the SECOND conv of a chain (weight and bias) by k multiplies every
    activation behind it by ~k (the later biases become negligible), dividing the chain's LAST (activation-free) conv weight
    by k brings the output back.  Head: cnn1 * k, cnn3 / k.  IFBlock i: conv0.1 * k, lastconv * amp / k.  (The first conv of
    each chain reads bounded data -- frames, warped frames, flow -- and keeps ordinary weights.)"""
    net.encode.cnn1.weight.mul_(k); net.encode.cnn1.bias.mul_(k)
    net.encode.cnn3.weight.div_(k)
    for i in range(5):
        blk = getattr(net, f"block{i}")
        blk.conv0[1][0].weight.mul_(k); blk.conv0[1][0].bias.mul_(k)
        blk.lastconv[0].weight.mul_(amp / k)


def case_stress():
    """Trained-weight-like stress fixtures recorded from the unmodified reference (fp32, CPU):
    (i) "range": conv activations of the Head and of every IFBlock reach >= 3e4 (beyond what fp16 storage can follow);
    (ii) "bigflow": lastconv * 20 drives |flow| to 50-200 px through the full net (warp border clamps everywhere)."""
    arrs = {}
    torch.manual_seed(1234)
    m = Model()
    m.eval()
    apply_range_stress(m.flownet)
    peak = {}
    hooks = []
    for name, mod in m.flownet.named_modules():
        if name.endswith("convblock.7") or name == "encode.cnn2":
            hooks.append(mod.register_forward_hook(lambda _m, _i, o, name=name: peak.__setitem__(name, max(peak.get(name, 0.), float(o.abs().max())))))
    img0, img1 = rand_frames(71, 1, 3, 128, 192, smooth=True)
    out = m.inference(img0, img1, 0.5)
    fl, _, _ = m.flownet(torch.cat((img0, img1), 1), 0.5, [16, 8, 4, 2, 1])
    for h in hooks:
        h.remove()
    print("range stress: peak |activation|", {k: f"{v:.3g}" for k, v in peak.items()}, "|flow|max", float(fl[-1].abs().max()))
    assert min(peak.values()) >= 3e4, peak
    arrs["r_img0"], arrs["r_img1"], arrs["r_out"] = f32(img0), f32(img1), f32(out)
    arrs["r_peak_act"] = np.asarray([peak[k] for k in sorted(peak)], np.float32)
    arrs["r_flow_absmax"] = np.float32(fl[-1].abs().max())
    mb = build_model(1234, BIGFLOW_AMP)
    img0, img1 = rand_frames(72, 1, 3, 192, 320, smooth=True)
    out = mb.inference(img0, img1, 0.5)
    fl, _, _ = mb.flownet(torch.cat((img0, img1), 1), 0.5, [16, 8, 4, 2, 1])
    q = torch.quantile(fl[-1].abs().flatten()[::7], torch.tensor([0.5, 0.9]))
    print("bigflow stress: |flow| median / p90 / max", q.tolist(), float(fl[-1].abs().max()))
    arrs["f_img0"], arrs["f_img1"], arrs["f_out"] = f32(img0), f32(img1), f32(out)
    arrs["f_flow"] = f32(fl[-1]).astype(np.float16)
    arrs["f_flow_absmax"] = np.float32(fl[-1].abs().max())
    save("stress_cases", **arrs)


def case_eccv_ref_fp16():
    """How far the reference's OWN half-precision mode (inference_video.py --fp16: default tensor type Half, so weights,
    frames and the warp grid of model/warplayer.py are all fp16) lands from its fp32 result on the eccv_cases inputs.
    Recorded so that the fp16-mode parity tests can show the B200 path is at least as accurate as that mode."""
    from model.RIFE import Model as EccvModel
    import model.warplayer as wl
    torch.manual_seed(1234)
    m = EccvModel()
    m.eval()
    for b in (m.flownet.block0, m.flownet.block1, m.flownet.block2):
        b.lastconv.weight.data.mul_(AMP)
    g = np.load(os.path.join(OUT, "eccv_cases.npz"))
    res = {}
    for case, kw in (("a", dict(scale_list=[4, 2, 1])), ("b", dict(scale=0.5, scale_list=[4, 2, 1])),
                     ("c", dict(scale_list=[4, 2, 1]))):
        i0, i1 = torch.from_numpy(g[case + "_img0"]).float(), torch.from_numpy(g[case + "_img1"]).float()
        torch.set_default_dtype(torch.float32)
        m.flownet.float()
        wl.backwarp_tenGrid.clear()
        ref = m.inference(i0, i1, **kw)
        torch.set_default_dtype(torch.float16)
        m.flownet.half()
        wl.backwarp_tenGrid.clear()
        out = m.inference(i0.half(), i1.half(), **kw).float()
        torch.set_default_dtype(torch.float32)
        e = (out - ref).abs()
        ok = bool(torch.isfinite(out).all())       # case c overflows to NaN in the reference's fp16 mode
        res[case] = {"finite": ok, "max_abs": float(e.max()) if ok else None,
                     "psnr_db": float(-10 * torch.log10((e ** 2).mean())) if ok else None}
    with open(os.path.join(OUT, "eccv_ref_fp16.json"), "w") as f:
        json.dump({"what": "reference model/RIFE.py in its own fp16 mode (CPU, torch %s) vs its fp32 output" % torch.__version__,
                   "cases": res}, f, indent=1)
    print(res)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "eccv":
        case_eccv()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "scale_list":
        with torch.no_grad():
            case_scale_list()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "stress":
        with torch.no_grad():
            case_stress()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "scale_up":
        case_scale_up()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "scene":
        case_scene_ssim()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "eccv_2r":
        with torch.no_grad():
            case_eccv_2r()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "eccv_m":
        with torch.no_grad():
            case_eccv_m()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "eccv_fp16":
        with torch.no_grad():
            case_eccv_ref_fp16()
        sys.exit(0)
    weights_manifest()
    case_demo()
    case_full_structure()
    case_inference_variants()
    case_scale_up()
    case_warp()
    case_ops()
    case_eccv()
