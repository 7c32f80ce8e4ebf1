"""Pins the CPU oracles (oracle/rife_oracle.py numpy, oracle/torch_ref.py) against the golden vectors produced
by the unmodified reference (tests/golden/make_golden.py).  The reference itself has no tests."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import AMP, GOLDEN, load_golden, numpy_state_dict, seeded_model, stress_model, torch_state_dict
from oracle import rife_oracle as O
from oracle import torch_ref as T

TOL = 2e-5  # fp32 re-association noise between numpy/BLAS and ATen/oneDNN


def test_mirror_regenerates_reference_weights(pkg):
    man = json.load(open(os.path.join(GOLDEN, "weights_seed1234.json")))
    m = seeded_model(pkg)
    sd = m.flownet.state_dict()
    assert len(sd) == man["n_tensors"] == 158
    assert sum(v.numel() for v in sd.values()) == man["n_params"] == 5664776
    h = hashlib.sha256()
    for k, v in sd.items():
        assert list(v.shape) == man["tensors"][k]["shape"], k
        a = v.detach().cpu().numpy().astype(np.float32)
        h.update(k.encode())
        h.update(a.tobytes())
    assert h.hexdigest() == man["sha256"]
    # SURVEY 8a known answer
    assert abs(float(sd["block0.conv0.0.0.weight"].sum()) - (-1.70219052)) < 1e-5


def test_warp_cases():
    g = load_golden("warp_cases")
    for name in ("small", "large", "huge"):
        x, fl = g[f"{name}_x"], g[f"{name}_flow"]
        # scalar-formula base grid: 1-ulp linspace differences vs ATen's vectorised kernel (see O.warp)
        assert np.abs(O.warp(x, fl) - g[f"{name}_y"]).max() < 2e-5, name
        grids = (torch.linspace(-1.0, 1.0, x.shape[3]).numpy(), torch.linspace(-1.0, 1.0, x.shape[2]).numpy())
        assert np.abs(O.warp(x, fl, grids) - g[f"{name}_y"]).max() < 5e-7, name
        yt = T.warp(torch.from_numpy(g[f"{name}_x"]), torch.from_numpy(g[f"{name}_flow"])).numpy()
        assert np.abs(yt - g[f"{name}_y"]).max() < 1e-6, name


def test_linspace_matches_torch():
    for n in (2, 3, 17, 64, 448, 1920):
        assert np.abs(O.linspace_m1_1(n) - torch.linspace(-1.0, 1.0, n).numpy()).max() <= 1.2e-7


def test_elementary_ops_against_torch():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 5, 12, 20)).astype(np.float32)
    w = rng.standard_normal((7, 5, 3, 3)).astype(np.float32)
    b = rng.standard_normal(7).astype(np.float32)
    for stride in (1, 2):
        ref = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride, 1).numpy()
        assert np.abs(O.conv2d(x, w, b, stride, 1) - ref).max() < 1e-4
    wt = rng.standard_normal((5, 8, 4, 4)).astype(np.float32)
    bt = rng.standard_normal(8).astype(np.float32)
    ref = torch.nn.functional.conv_transpose2d(torch.from_numpy(x), torch.from_numpy(wt), torch.from_numpy(bt), 2, 1)
    assert np.abs(O.conv_transpose2d_k4s2p1(x, wt, bt) - ref.numpy()).max() < 1e-4
    assert np.array_equal(O.pixel_shuffle2(ref.numpy()), torch.nn.functional.pixel_shuffle(ref, 2).numpy())
    for sf in (0.5, 0.25, 0.125, 1.0, 2.0, 4.0, 16.0):
        xx = rng.standard_normal((1, 3, 16, 32)).astype(np.float32)
        ref = torch.nn.functional.interpolate(torch.from_numpy(xx), scale_factor=sf, mode="bilinear", align_corners=False)
        assert np.abs(O.interpolate_bilinear(xx, sf) - ref.numpy()).max() < 1e-6, sf


def test_ops_cases(pkg):
    g = load_golden("ops_cases")
    sd = numpy_state_dict(seeded_model(pkg))
    assert np.abs(O.head_forward(sd, g["head_x"]) - g["head_y"]).max() < TOL
    f, m, ft = O.ifblock_forward(sd, "block3.", g["blk3_x"], g["blk3_flow"], 2)
    assert np.abs(f - g["blk3_oflow"]).max() < TOL and np.abs(m - g["blk3_omask"]).max() < TOL
    assert np.abs(ft - g["blk3_ofeat"]).max() < TOL
    f, m, ft = O.ifblock_forward(sd, "block0.", g["blk0_x"], None, 16)
    assert np.abs(f - g["blk0_oflow"]).max() < TOL * 16 and np.abs(ft - g["blk0_ofeat"]).max() < TOL


def test_full_structure(pkg):
    g = load_golden("ifnet_full_64x128_amp")
    sd = numpy_state_dict(seeded_model(pkg, AMP))
    x = np.concatenate([g["img0"], g["img1"]], 1)
    flows, mask, merged = O.ifnet_forward(sd, x, 0.5, [16, 8, 4, 2, 1])
    for i in range(5):
        assert np.abs(flows[i] - g[f"flow{i}"]).max() < 2e-4, i   # |flow| reaches ~10 px
    for i in range(4):
        assert np.abs(merged[i][0] - g[f"merged{i}_0"]).max() < 1e-4
        assert np.abs(merged[i][1] - g[f"merged{i}_1"]).max() < 1e-4
    assert np.abs(mask - g["mask"]).max() < 1e-4
    assert np.abs(merged[4] - g["merged4"]).max() < 1e-4


@pytest.mark.parametrize("impl", ["numpy", "torch"])
def test_inference_variants(pkg, impl):
    g = load_golden("inference_variants")

    def run(amp, a, b, t=0.5, scale=1.0):
        m = seeded_model(pkg, amp)
        if impl == "numpy":
            return O.model_inference(numpy_state_dict(m), a, b, t, scale)
        tt = torch.from_numpy(np.asarray(t, np.float32)) if np.ndim(t) else t
        return T.model_inference(torch_state_dict(m), torch.from_numpy(a), torch.from_numpy(b), tt, scale).numpy()

    assert np.abs(run(AMP, g["a_img0"], g["a_img1"], 0.3) - g["a_out"]).max() < 1e-4
    assert np.abs(run(AMP, g["b_img0"], g["b_img1"], 0.5, 0.5) - g["b_out"]).max() < 1e-4
    assert np.abs(run(1.0, g["c_img0"], g["c_img1"]) - g["c_out"]).max() < 1e-4
    assert np.abs(run(AMP, g["d_img0"], g["d_img1"], g["d_t"]) - g["d_out"]).max() < 1e-4
    assert np.abs(run(1.0, g["e_img0"], g["e_img1"]) - g["e_out"]).max() < 1e-4


@pytest.mark.parametrize("impl", ["numpy", "torch"])
def test_scale_above_one(pkg, impl):
    """--scale 2 / 4: the last block(s) run above full resolution (scale_list entries < 1)."""
    g = load_golden("scale_up_cases")
    m = seeded_model(pkg, AMP)
    for tag, scale in (("s2", 2.0), ("s4", 4.0)):
        a, b = g[tag + "_img0"], g[tag + "_img1"]
        if impl == "numpy":
            out = O.model_inference(numpy_state_dict(m), a, b, 0.5, scale)
        else:
            out = T.model_inference(torch_state_dict(m), torch.from_numpy(a), torch.from_numpy(b), 0.5, scale).numpy()
        assert np.abs(out - g[tag + "_out"]).max() < 1e-4, tag


def test_demo_config1(pkg):
    """BASELINE config 1: demo/I0_0.png + I0_1.png, exp=1, scale=1.0, fp32 (known answer of SURVEY 8a)."""
    g = load_golden("demo_256x448")
    img0 = (g["img0_u8"].transpose(2, 0, 1)[None].astype(np.float32) / np.float32(255.))
    img1 = (g["img1_u8"].transpose(2, 0, 1)[None].astype(np.float32) / np.float32(255.))
    out = O.model_inference(numpy_state_dict(seeded_model(pkg)), img0, img1)
    assert abs(float(g["out"].mean()) - 0.34859153) < 1e-6
    assert np.abs(out - g["out"]).max() < 1e-4
    out_amp = T.model_inference(torch_state_dict(seeded_model(pkg, AMP)), torch.from_numpy(img0), torch.from_numpy(img1)).numpy()
    assert np.abs(out_amp - g["out_amp"]).max() < 1e-4


def test_stress_cases_torch_oracle(pkg):
    """The trained-weight-like stress fixtures (tests/golden/make_golden.py:case_stress, recorded from the unmodified
    reference): conv activations up to 5.8e5 ("range") and |flow| up to 240 px ("bigflow").  The torch-CPU restatement
    reproduces both; the mirror's weight transformation is therefore the fixture's."""
    g = load_golden("stress_cases")
    assert float(g["r_peak_act"].min()) >= 3e4 and float(g["r_peak_act"].max()) > 65504.0
    assert float(g["f_flow_absmax"]) > 200.0
    m = stress_model(pkg, "range", "fp32")
    out = T.model_inference(torch_state_dict(m), torch.from_numpy(g["r_img0"]), torch.from_numpy(g["r_img1"]), 0.5).numpy()
    assert np.abs(out - g["r_out"]).max() < 1e-4
    m = stress_model(pkg, "bigflow", "fp32")
    out = T.model_inference(torch_state_dict(m), torch.from_numpy(g["f_img0"]), torch.from_numpy(g["f_img1"]), 0.5).numpy()
    assert np.abs(out - g["f_out"]).max() < 1e-4


def test_arbitrary_scale_lists_torch_oracle(pkg):
    """scale lists other than the halving pyramid (tests/golden/make_golden.py:case_scale_list)."""
    g = load_golden("scale_list_cases")
    m = seeded_model(pkg, AMP)
    x = torch.cat((torch.from_numpy(g["img0"]), torch.from_numpy(g["img1"])), 1)
    for name in "abcd":
        sl = [float(v) for v in g[name + "_list"]]
        flows, mask, merged = T.ifnet_forward(torch_state_dict(m), x, 0.4, sl)
        assert np.abs(flows[4].numpy() - g[name + "_flow4"]).max() < 1e-4, name
        assert np.abs(merged[4].numpy() - g[name + "_merged"]).max() < 1e-5, name
