"""Drop-in for the reference's `train_log/RIFE_HDv3.py` `Model` facade (inference side)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .ifnet import IFNet

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class Model:
    """Same surface as RIFE_HDv3.py:15-60: Model(), .flownet, .version, .train/.eval/.device/.load_model/
    .save_model/.inference.  `update` (training, broken in the reference) is out of scope."""

    def __init__(self, local_rank=-1, precision="auto"):
        self.flownet = IFNet(precision=precision)
        self.device()
        self.version = 4.25
        if local_rank != -1:
            raise NotImplementedError("DDP training wrapper (RIFE_HDv3.py:24-25) is out of scope: inference only")

    def train(self):
        self.flownet.train()

    def eval(self):
        self.flownet.eval()

    def device(self):
        self.flownet.to(device)

    def load_model(self, path, rank=0):
        """RIFE_HDv3.py:36-50: rank -1 keeps only 'module.'-prefixed keys and strips the prefix; non-strict."""
        def convert(param):
            if rank == -1:
                return {k.replace("module.", ""): v for k, v in param.items() if "module." in k}
            return param

        if rank <= 0:
            sd = torch.load("{}/flownet.pkl".format(path), map_location="cpu" if not torch.cuda.is_available() else None)
            self.flownet.load_state_dict(convert(sd), False)

    def save_model(self, path, rank=0):
        if rank == 0:
            torch.save(self.flownet.state_dict(), "{}/flownet.pkl".format(path))

    def inference(self, img0, img1, timestep=0.5, scale=1.0, pad=False):
        """RIFE_HDv3.py:56-60.  Positional order (img0, img1, timestep, scale) is the reference's.
        Extensions: pad=True accepts any H, W (the zero-pad to 64/scale and the crop the reference's callers do around
        this call happen inside the first / last kernel); uint8 tensors are taken as HWC (B,H,W,C) frames."""
        return self.flownet.run(img0, img1, timestep, scale, want_all=False, pad=pad)

    def inference_multi(self, img0, img1, timesteps=(0.25, 0.5, 0.75), scale=1.0, pad=False):
        """Arbitrary-timestep nX interpolation (benchmark/HD_multi_4X.py:47 calls inference once per timestep): all
        `timesteps` of every pair in one call, Head features computed once per frame.  Returns (P, T, C, H, W)."""
        return self.flownet.run_multi(img0, img1, timesteps, scale, pad)

    def inference_sequence(self, frames, timesteps=(0.5,), scale=1.0, pad=False):
        """Video order (inference_video.py:215-260 walks consecutive frames): `frames` holds N consecutive frames, the
        result the (N-1, T, ...) in-between frames; each frame's Head features are computed once and shared by the two
        pairs it belongs to."""
        return self.flownet.run_multi(frames, None, timesteps, scale, pad)

    def interpolate_recursive(self, img0, img1, exp=1, timestep=0.5, scale=1.0, pad=False):
        """The midpoint-of-midpoints loop of inference_img.py:95-103 (and make_inference, inference_video.py:178-188,
        for n = 2**exp - 1): `exp` rounds of inserting the in-between frame of every adjacent pair.  The pairs of one
        round are independent, so each round is ONE batched inference call (2**k pairs x B) instead of 2**k calls.
        Returns the frame list [img0, ..., img1] (2**exp + 1 entries, each shaped like img0)."""
        frames = [img0, img1]
        b = img0.shape[0]
        for _ in range(int(exp)):
            left = torch.cat(frames[:-1], 0)
            right = torch.cat(frames[1:], 0)
            mids = self.flownet.run(left, right, timestep, scale, want_all=False, pad=pad)
            nxt = []
            for j in range(len(frames) - 1):
                nxt.append(frames[j])
                nxt.append(mids[j * b:(j + 1) * b])
            nxt.append(frames[-1])
            frames = nxt
        return frames

    def inference_host(self, img0_host, img1_host, out_host, timestep=0.5, scale=1.0, stream=None, slot=0, streams=None,
                       device=None):
        """End-to-end call on HOST (ideally pinned) tensors -- NCHW float16/float32, or HWC uint8 frames exactly as a
        video decoder delivers them (any H, W): H2D copies, inference and the D2H copy of the result are enqueued on
        `stream` (default: current) by rife_inference_host; returns without syncing.
        `img1_host=None` selects video order (inference_video.py:215-260): `img0_host` holds N + 1 consecutive frames,
        `out_host` receives the N in-between frames, and every frame crosses the host link (and the Head) once.
        `slot` selects an independent plan (workspace + staging buffers).  With `streams=(h2d, compute, d2h)` the call
        becomes one stage set of a software pipeline (rife_inference_host_pipelined): rotate two or three slots over the
        same three streams and the PCIe copies of one call overlap the kernels of its neighbours, while the kernels
        of consecutive calls run back to back on the single compute stream.
        `device`: the CUDA device to run on (default: the current device) -- one handle set per device, no module global."""
        net = self.flownet
        seq = img1_host is None
        for name, t in (("img0_host", img0_host), ("img1_host", img1_host), ("out_host", out_host)):
            if t is None and name == "img1_host":
                continue
            if not torch.is_tensor(t) or t.is_cuda:
                raise ValueError(f"inference_host: {name} must be a host (CPU) tensor")
            if not t.is_contiguous():
                raise ValueError(f"inference_host: {name} must be contiguous (got strides {t.stride()})")
        if img0_host.dtype not in (torch.float32, torch.float16, torch.uint8):
            raise TypeError("inference_host: float32 / float16 NCHW frames or uint8 HWC frames required")
        if img0_host.dim() != 4:
            raise ValueError("inference_host: frames must be 4-D (B,C,H,W) float or (B,H,W,C) uint8")
        if not seq and (img1_host.shape != img0_host.shape or img1_host.dtype != img0_host.dtype):
            raise ValueError("inference_host: img0_host and img1_host must have the same shape and dtype")
        lib = _lib.load()
        u8 = img0_host.dtype == torch.uint8
        if u8:
            n, h, w, c = img0_host.shape
        else:
            n, c, h, w = img0_host.shape
        b = n - 1 if seq else n
        if b < 1:
            raise ValueError("inference_host: video order needs at least two frames")
        if out_host.dtype != img0_host.dtype or out_host.numel() != b * c * h * w:
            raise ValueError(f"inference_host: out_host must hold {b} frames shaped like the inputs in {img0_host.dtype} "
                             f"(got {tuple(out_host.shape)} {out_host.dtype})")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("eccv2022-rife_b200: a CUDA device is required -- the B200 hot path has no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        prec = net._resolve_precision(torch.float16 if u8 else img0_host.dtype)
        pl = net.plan_for(dev, prec, b, c, h, w, scale, slot, True, 1, seq)
        io = net._io(img0_host.dtype)
        ts = net._timesteps(timestep, b)
        p1 = None if seq else img1_host.data_ptr()
        with torch.cuda.device(dev):
            if streams is not None:
                s_in, s_c, s_out = (x.cuda_stream for x in streams)
                _lib.check(lib.rife_inference_host_pipelined(pl.handle, img0_host.data_ptr(), p1, ts,
                                                             out_host.data_ptr(), io, s_in, s_c, s_out))
                return out_host
            st = (stream or torch.cuda.current_stream()).cuda_stream
            _lib.check(lib.rife_inference_host(pl.handle, img0_host.data_ptr(), p1, ts, out_host.data_ptr(), io, st))
        return out_host
