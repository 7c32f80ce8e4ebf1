"""ECCV-paper network (BASELINE config 5): oracles pinned to goldens recorded from the unmodified reference."""
import hashlib
import json
import os

import numpy as np
import torch

from conftest import AMP, GOLDEN, load_golden, seeded_eccv, seeded_eccv_2r, seeded_eccv_m
from oracle import rife_oracle as O
from oracle import torch_ref as T


def test_eccv_mirror_regenerates_reference_weights(pkg):
    man = json.load(open(os.path.join(GOLDEN, "weights_eccv_seed1234.json")))
    m = seeded_eccv(pkg)
    sd = m._sha_state
    assert len(sd) == man["n_tensors"] == 190 and sum(v.numel() for v in sd.values()) == man["n_params"] == 10708215
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.numpy().astype(np.float32).tobytes())
    assert h.hexdigest() == man["sha256"]


def test_eccv_oracles_against_golden(pkg):
    g = load_golden("eccv_cases")
    m = seeded_eccv(pkg)
    sdt = {k: v.detach().cpu().float() for k, v in m.flownet.state_dict().items()}
    sdn = {k: v.numpy() for k, v in sdt.items()}
    x = np.concatenate([g["a_img0"], g["a_img1"]], 1)
    for name, (flows, mask, merged) in {"numpy": O.eccv_ifnet_forward(sdn, x, [4, 2, 1]),
                                        "torch": T.eccv_ifnet_forward(sdt, torch.from_numpy(x), [4, 2, 1])}.items():
        for i in range(3):
            assert np.abs(np.asarray(flows[i]) - g[f"a_flow{i}"]).max() < 5e-4, (name, i)
            assert np.abs(np.asarray(merged[i]) - g[f"a_merged{i}"]).max() < 1e-4, (name, i)
        assert np.abs(np.asarray(mask) - g["a_mask"]).max() < 1e-4, name
    out = T.eccv_model_inference(sdt, torch.from_numpy(g["b_img0"]), torch.from_numpy(g["b_img1"]), 0.5).numpy()
    assert np.abs(out - g["b_out"]).max() < 1e-4


def test_eccv_m_mirror_weights_and_oracle(pkg):
    """model/IFNet_m.py (Model(arbitrary=True)): the mirror regenerates the reference's seeded weights and the torch
    restatement reproduces the reference's outputs at several timesteps."""
    man = json.load(open(os.path.join(GOLDEN, "weights_eccv_m_seed1234.json")))
    m = seeded_eccv_m(pkg)
    h = hashlib.sha256()
    for k, v in m._sha_state.items():
        h.update(k.encode())
        h.update(v.numpy().astype(np.float32).tobytes())
    assert h.hexdigest() == man["sha256"] and len(m._sha_state) == man["n_tensors"]
    g = load_golden("eccv_m_cases")
    sdt = {k: v.detach().cpu().float() for k, v in m.flownet.state_dict().items()}
    x = torch.from_numpy(np.concatenate([g["img0"], g["img1"]], 1))
    for t in (0.25, 0.5, 0.8):
        flows, mask, merged = T.eccv_ifnet_forward(sdt, x, [4, 2, 1], timestep=t)
        assert np.abs(flows[2].numpy() - g[f"flow2_t{t}"]).max() < 5e-4
        assert np.abs(mask.numpy() - g[f"mask_t{t}"]).max() < 1e-4
        assert np.abs(merged[2].numpy() - g[f"out_t{t}"]).max() < 1e-4
    out = T.eccv_model_inference(sdt, torch.from_numpy(g["img0"]), torch.from_numpy(g["img1"]), 0.5, timestep=0.25).numpy()
    assert np.abs(out - g["out_scale05_t0.25"]).max() < 1e-4


def test_eccv_2r_mirror_weights_and_oracle(pkg):
    """model/IFNet_2R.py + model/refine_2R.py (the stride variant): same 190 tensors and default init as model/IFNet.py;
    the torch restatement with two_r=True reproduces the outputs recorded from the reference's IFNet_2R.forward."""
    man = json.load(open(os.path.join(GOLDEN, "weights_eccv_seed1234.json")))
    net = seeded_eccv_2r(pkg)
    h = hashlib.sha256()
    for k, v in net._sha_state.items():
        h.update(k.encode())
        h.update(v.numpy().astype(np.float32).tobytes())
    assert h.hexdigest() == man["sha256"] and len(net._sha_state) == man["n_tensors"]
    assert net.block0.conv0[0][0].stride == (1, 1) and net.unet.conv.stride == (2, 2) and net.unet.down0.conv1[0].stride == (1, 1)
    g = load_golden("eccv_2r_cases")
    sdt = {k: v.detach().cpu().float() for k, v in net.state_dict().items()}
    x = torch.from_numpy(np.concatenate([g["a_img0"], g["a_img1"]], 1))
    flows, mask, merged = T.eccv_ifnet_forward(sdt, x, [4, 2, 1], two_r=True)
    for i in range(3):
        assert np.abs(flows[i].numpy() - g[f"a_flow{i}"]).max() < 5e-4, i
    assert np.abs(mask.numpy() - g["a_mask"]).max() < 1e-4
    assert np.abs(merged[0].numpy() - g["a_merged0"]).max() < 1e-4
    assert np.abs(merged[2].numpy() - g["a_merged2"]).max() < 1e-4
    xb = torch.from_numpy(np.concatenate([g["b_img0"], g["b_img1"]], 1))
    assert np.abs(T.eccv_ifnet_forward(sdt, xb, [8, 4, 2], two_r=True)[2][2].numpy() - g["b_out"]).max() < 1e-4
