import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
AMP = 4.0  # tests/golden/make_golden.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("eccv2022-rife_b200")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


_models = {}


def seeded_model(pkg, amp=1.0, precision="auto"):
    """The host mirror under torch.manual_seed(1234): regenerates the reference's default-init weights
    (checked against tests/golden/weights_seed1234.json); `amp` scales every lastconv weight."""
    import torch
    key = (amp, precision)
    if key not in _models:
        torch.manual_seed(1234)
        m = pkg.Model(precision=precision)
        m.eval()
        if amp != 1.0:
            with torch.no_grad():
                for i in range(5):
                    getattr(m.flownet, f"block{i}").lastconv[0].weight.mul_(amp)
            m.flownet.invalidate()
        _models[key] = m
    return _models[key]


RANGE_K, BIGFLOW_AMP = 2.0e5, 20.0   # tests/golden/make_golden.py: case_stress


def stress_model(pkg, kind, precision="auto"):
    """The mirror with the weight transformations of tests/golden/make_golden.py:case_stress.
    "range": conv activations of the Head and of every IFBlock * 2e5 (>= 3e4 everywhere, > the fp16 maximum in blocks
    0-4), the activation-free last conv of each chain / 2e5; "bigflow": every lastconv * 20 (|flow| up to 240 px)."""
    import torch
    key = ("stress", kind, precision)
    if key not in _models:
        torch.manual_seed(1234)
        m = pkg.Model(precision=precision)
        m.eval()
        net = m.flownet
        with torch.no_grad():
            if kind == "range":
                net.encode.cnn1.weight.mul_(RANGE_K); net.encode.cnn1.bias.mul_(RANGE_K)
                net.encode.cnn3.weight.div_(RANGE_K)
                for i in range(5):
                    blk = getattr(net, f"block{i}")
                    blk.conv0[1][0].weight.mul_(RANGE_K); blk.conv0[1][0].bias.mul_(RANGE_K)
                    blk.lastconv[0].weight.mul_(AMP / RANGE_K)
            elif kind == "bigflow":
                for i in range(5):
                    getattr(net, f"block{i}").lastconv[0].weight.mul_(BIGFLOW_AMP)
            else:
                raise ValueError(kind)
        net.invalidate()
        _models[key] = m
    return _models[key]


def numpy_state_dict(model):
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in model.flownet.state_dict().items()}


def torch_state_dict(model):
    return {k: v.detach().cpu().float() for k, v in model.flownet.state_dict().items()}


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)


_eccv = {}


_eccv_m = {}


def seeded_eccv_m(pkg, amp=AMP):
    """Arbitrary-timestep ECCV model mirror (Model(arbitrary=True) = IFNet_m), seeded like tests/golden/make_golden.py:
    case_eccv_m (sha256 in tests/golden/weights_eccv_m_seed1234.json)."""
    import torch
    if amp not in _eccv_m:
        torch.manual_seed(1234)
        m = pkg.eccv.Model(arbitrary=True)
        m.eval()
        m._sha_state = {k: v.detach().cpu().clone() for k, v in m.flownet.state_dict().items()}
        if amp != 1.0:
            with torch.no_grad():
                for b in (m.flownet.block0, m.flownet.block1, m.flownet.block2):
                    b.lastconv.weight.mul_(amp)
            m.flownet.invalidate()
        _eccv_m[amp] = m
    return _eccv_m[amp]


_eccv_2r = {}


def seeded_eccv_2r(pkg, amp=AMP):
    """IFNet_2R mirror (model/IFNet_2R.py + model/refine_2R.py), seeded like tests/golden/make_golden.py:case_eccv_2r.
    Same tensors and default init as the ECCV net (sha256 in tests/golden/weights_eccv_seed1234.json)."""
    import torch
    if amp not in _eccv_2r:
        torch.manual_seed(1234)
        net = pkg.eccv.IFNet_2R().to("cuda" if torch.cuda.is_available() else "cpu")
        net.eval()
        net._sha_state = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
        if amp != 1.0:
            with torch.no_grad():
                for b in (net.block0, net.block1, net.block2):
                    b.lastconv.weight.mul_(amp)
            net.invalidate()
        _eccv_2r[amp] = net
    return _eccv_2r[amp]


def seeded_eccv(pkg, amp=AMP):
    """ECCV-paper model mirror under torch.manual_seed(1234) (sha256 in tests/golden/weights_eccv_seed1234.json);
    `amp` scales block{0,1,2}.lastconv.weight as tests/golden/make_golden.py:case_eccv does."""
    import torch
    if amp not in _eccv:
        torch.manual_seed(1234)
        m = pkg.eccv.Model()
        m.eval()
        m._sha_state = {k: v.detach().cpu().clone() for k, v in m.flownet.state_dict().items()}
        if amp != 1.0:
            with torch.no_grad():
                for b in (m.flownet.block0, m.flownet.block1, m.flownet.block2):
                    b.lastconv.weight.mul_(amp)
            m.flownet.invalidate()
        _eccv[amp] = m
    return _eccv[amp]
