"""ECCV-paper network on the GPU (generic NHWC layer API, fp32) against the reference's golden vectors."""
import numpy as np
import pytest
import torch

from conftest import load_golden, seeded_eccv
from oracle import torch_ref as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_eccv_ifnet_forward_structure(pkg):
    g = load_golden("eccv_cases")
    m = seeded_eccv(pkg)
    x = torch.cat((cu(g["a_img0"]), cu(g["a_img1"])), 1)
    flows, mask, merged, ft, mt, ld = m.flownet(x, [4, 2, 1])
    assert ft is None and mt is None and ld == 0 and len(flows) == 3 and len(merged) == 3
    for i in range(3):
        assert np.abs(flows[i].cpu().numpy() - g[f"a_flow{i}"]).max() < 1e-3, i      # |flow| up to ~28 px
        assert np.abs(merged[i].cpu().numpy() - g[f"a_merged{i}"]).max() < 1e-4, i
    assert np.abs(mask.cpu().numpy() - g["a_mask"]).max() < 1e-4


def test_eccv_inference_scale_and_vimeo_shape(pkg):
    g = load_golden("eccv_cases")
    m = seeded_eccv(pkg)
    out = m.inference(cu(g["b_img0"]), cu(g["b_img1"]), scale=0.5, scale_list=[4, 2, 1]).cpu().numpy()
    assert np.abs(out - g["b_out"]).max() < 1e-4
    a, b = cu(g["c_img0"].astype(np.float32)), cu(g["c_img1"].astype(np.float32))
    out = m.inference(a, b, scale_list=[4, 2, 1]).cpu().numpy()                      # 256x448: benchmark/Vimeo90K.py shape
    assert np.abs(out - g["c_out"].astype(np.float32)).max() < 1e-3                    # golden stored as fp16


def test_eccv_batch_and_tta(pkg):
    m = seeded_eccv(pkg)
    rng = np.random.default_rng(8)
    a, b = cu(rng.random((3, 3, 64, 64), dtype=np.float32)), cu(rng.random((3, 3, 64, 64), dtype=np.float32))
    full = m.inference(a, b)
    for i in range(3):
        assert torch.equal(full[i:i + 1], m.inference(a[i:i + 1], b[i:i + 1]))
    sd = {k: v.detach().cpu().float() for k, v in m.flownet.state_dict().items()}
    ref = T.eccv_model_inference(sd, a.cpu(), b.cpu()).numpy()
    assert np.abs(full.cpu().numpy() - ref).max() < 1e-4
    tta = m.inference(a, b, TTA=True)
    assert tta.shape == full.shape and float(tta.min()) >= 0 and float(tta.max()) <= 1
    with pytest.raises(RuntimeError):
        m.inference(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 64, 64))               # CPU tensors
    with pytest.raises(RuntimeError, match="multiples of"):
        m.inference(torch.zeros(1, 3, 40, 64, device=DEV), torch.zeros(1, 3, 40, 64, device=DEV))
