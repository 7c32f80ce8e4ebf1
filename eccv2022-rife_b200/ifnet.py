"""Host-side mirror of the reference's `train_log/IFNet_HDv3.py` (RIFE v4.25 HDv3 IFNet).

The nn.Modules here are *parameter containers*: they register exactly the reference's 158
state-dict tensors (same keys, shapes and construction order, so the default init under a given
torch seed and `load_state_dict` of `train_log/flownet.pkl` behave identically) but never run a
PyTorch conv.  `IFNet.forward` hands raw device pointers to librife_b200.so, which runs the whole
pipeline with hand-written sm_100a kernels (see csrc/ and DESIGN.md).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib
from .warplayer import base_grids

_BLOCK_C = (192, 128, 96, 64, 32)     # IFNet_HDv3.py:95-99
_BLOCK_IN = (7 + 8, 8 + 4 + 8 + 8, 8 + 4 + 8 + 8, 8 + 4 + 8 + 8, 8 + 4 + 8 + 8)


def _conv_lrelu(cin, cout, stride):
    # key layout "<name>.0.{weight,bias}" of the reference's conv() helper (IFNet_HDv3.py:9-14)
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1, bias=True), nn.LeakyReLU(0.2, True))


class Head(nn.Module):
    """Parameters of IFNet_HDv3.py:24-31 (cnn0..cnn3)."""

    def __init__(self):
        super().__init__()
        self.cnn0 = nn.Conv2d(3, 16, 3, 2, 1)
        self.cnn1 = nn.Conv2d(16, 16, 3, 1, 1)
        self.cnn2 = nn.Conv2d(16, 16, 3, 1, 1)
        self.cnn3 = nn.ConvTranspose2d(16, 4, 4, 2, 1)


class ResConv(nn.Module):
    """Parameters of IFNet_HDv3.py:45-51 (conv, beta)."""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 1, 1)
        self.beta = nn.Parameter(torch.ones((1, c, 1, 1)), requires_grad=True)


class IFBlock(nn.Module):
    """Parameters of IFNet_HDv3.py:57-76 (conv0, convblock, lastconv)."""

    def __init__(self, in_planes, c=64):
        super().__init__()
        self.conv0 = nn.Sequential(_conv_lrelu(in_planes, c // 2, 2), _conv_lrelu(c // 2, c, 2))
        self.convblock = nn.Sequential(*[ResConv(c) for _ in range(8)])
        self.lastconv = nn.Sequential(nn.ConvTranspose2d(c, 4 * 13, 4, 2, 1), nn.PixelShuffle(2))


def _precision_from_env(default):
    v = os.environ.get("RIFE_B200_PRECISION", default).lower()
    if v not in ("fp32", "fp16", "bf16", "auto"):
        raise ValueError("RIFE_B200_PRECISION must be fp32, fp16, bf16 or auto")
    return v


class _Plan:
    __slots__ = ("handle", "workspace", "key", "frame")


class IFNet(nn.Module):
    """Drop-in for `train_log.IFNet_HDv3.IFNet`.

    precision: "fp32" (fp32 storage + FMA convs, <=1e-4 vs the reference), "fp16" (fp16 NHWC
    activations, tcgen05 tensor-core convs, fp32 accumulate/coordinates), "bf16" (as fp16 with the conv
    activations inside the Head / IFBlocks in bf16: fp32 range, cannot overflow) or "auto" (follows the input
    dtype: float16 -> fp16 mode, bfloat16 -> bf16 mode, float32 -> fp32 mode).  Env override: RIFE_B200_PRECISION.
    bfloat16 frame tensors are accepted in every 16-bit mode: they cross the boundary as float16 (exact for
    values in [0, 1]) and the result is cast back.
    """

    def __init__(self, precision="auto"):
        super().__init__()
        for i in range(5):                                    # same order as IFNet_HDv3.py:95-100
            setattr(self, f"block{i}", IFBlock(_BLOCK_IN[i], c=_BLOCK_C[i]))
        self.encode = Head()
        self.precision = _precision_from_env(precision)
        self._weights = {}    # (device index, precision) -> (handle, generation)
        self._plans = {}
        self._generation = 0

    # ------------------------------------------------------------------ cache invalidation
    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._generation += 1
        return r

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._generation += 1
        return r

    def invalidate(self):
        """Call after mutating parameters in place (the packed device copy is rebuilt on next use)."""
        self._generation += 1

    def _param_stamp(self):
        return self._generation, sum(p._version for p in self.parameters())

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def release(self):
        lib = _lib._lib
        if lib is None:
            return
        for pl in self._plans.values():
            lib.rife_plan_destroy(pl.handle)
        self._plans.clear()
        for h, _ in self._weights.values():
            lib.rife_weights_destroy(h)
        self._weights.clear()

    # ------------------------------------------------------------------ engine plumbing
    def _resolve_precision(self, dtype):
        p = self.precision
        if p == "auto":
            p = "fp16" if dtype == torch.float16 else ("bf16" if dtype == torch.bfloat16 else "fp32")
        return {"fp16": _lib.PRECISION_FP16, "bf16": _lib.PRECISION_BF16}.get(p, _lib.PRECISION_FP32)

    def _get_weights(self, device, prec):
        lib = _lib.load()
        key = (device.index, prec)
        stamp = self._param_stamp()
        ent = self._weights.get(key)
        if ent is not None and ent[1] == stamp:
            return ent[0]
        if ent is not None:   # stale: plans bound to it go too
            for k in [k for k in self._plans if k[0] == key]:
                lib.rife_plan_destroy(self._plans.pop(k).handle)
            lib.rife_weights_destroy(ent[0])
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in self.state_dict().items()}
        n = len(sd)
        names = (C.c_char_p * n)(*[k.encode() for k in sd])
        ptrs = (C.c_void_p * n)(*[v.data_ptr() for v in sd.values()])
        numel = (C.c_int64 * n)(*[v.numel() for v in sd.values()])
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.rife_weights_create(C.byref(h), n, names, ptrs, numel, prec))
        self._weights[key] = (h, stamp)
        return h

    def _get_plan(self, device, prec, b, c, h, w, scale, slot=0, nt=1, seq=False):
        """b frame pairs x nt timesteps per pair; seq: the pairs are the b consecutive pairs of one run of b+1 frames."""
        lib = _lib.load()
        wh = self._get_weights(device, prec)
        scale = tuple(float(x) for x in scale) if isinstance(scale, (tuple, list)) else float(scale)
        key = ((device.index, prec), b, c, h, w, scale, slot, nt, bool(seq))
        pl = self._plans.pop(key, None)
        if pl is not None:
            self._plans[key] = pl                  # most recently used goes last
            return pl
        with torch.cuda.device(device):
            if isinstance(scale, tuple):   # arbitrary scale_list (IFNet.forward)
                sl = (C.c_float * 5)(*scale)
                nbytes = lib.rife_plan_workspace_bytes_list(b, nt, int(seq), c, h, w, sl, prec)
            else:
                nbytes = lib.rife_plan_workspace_bytes_ex(b, nt, int(seq), c, h, w, float(scale), prec)
            if nbytes == 0:
                _lib.check(-2)
            pl = _Plan()
            pl.workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)
            gx, gy, _, _ = base_grids(h, w, device)
            hd = C.c_void_p()
            if isinstance(scale, tuple):
                _lib.check(lib.rife_plan_create_list(C.byref(hd), wh, b, nt, int(seq), c, h, w, sl,
                                                     pl.workspace.data_ptr(), nbytes, gx.data_ptr(), gy.data_ptr()))
            else:
                _lib.check(lib.rife_plan_create_ex(C.byref(hd), wh, b, nt, int(seq), c, h, w, float(scale),
                                                   pl.workspace.data_ptr(), nbytes, gx.data_ptr(), gy.data_ptr()))
            pl.handle = hd
        self._plans[key] = pl
        self._evict_plans(device)
        return pl

    def _evict_plans(self, device):
        """Least-recently-used eviction: a plan owns a workspace of ~0.5 GB per 1080p pair, and callers such as
        interpolate_recursive create one per batch size.  At most RIFE_B200_MAX_PLANS (default 8) plans stay alive;
        an evicted plan is destroyed after the device has drained (its kernels may still be running)."""
        limit = max(1, int(os.environ.get("RIFE_B200_MAX_PLANS", "8")))
        if len(self._plans) <= limit:
            return
        torch.cuda.synchronize(device)
        lib = _lib.load()
        while len(self._plans) > limit:
            k = next(iter(self._plans))            # dicts keep insertion order; _get_plan re-inserts on every hit
            lib.rife_plan_destroy(self._plans.pop(k).handle)

    @staticmethod
    def _scale_from_list(scale_list):
        if len(scale_list) < 5:
            raise ValueError("IFNet.forward indexes scale_list[0..4] (IFNet_HDv3.py:139-148): pass 5 entries")
        s4 = float(scale_list[4])
        sl = [float(s) for s in scale_list[:5]]
        if sl == [16 * s4, 8 * s4, 4 * s4, 2 * s4, s4] and s4 in (0.25, 0.5, 1.0, 2.0, 4.0):
            return 1.0 / s4                        # the pyramid of RIFE_HDv3.py:58: fast-path kernels
        for s in sl:
            m, _ = math.frexp(s)
            if not (0.25 <= s <= 64.0) or m != 0.5:
                raise NotImplementedError(
                    f"eccv2022-rife_b200 takes scale_list entries that are powers of two in [1/4, 64] (every list "
                    f"[16,8,4,2,1]/scale of RIFE_HDv3.py:58 / inference_video.py:92 is); got {list(scale_list)}")
        return tuple(sl)                           # any other list: generic pixel kernels (rife_plan_create_list)

    @staticmethod
    def _timesteps(timestep, b):
        if torch.is_tensor(timestep):
            t = timestep.detach().to("cpu", torch.float32).reshape(-1)
            if t.numel() == 1:
                t = t.repeat(b)
            if t.numel() != b:
                raise ValueError("tensor timestep must have shape (B,1,1,1) (IFNet_HDv3.py:127)")
            vals = t.tolist()
        else:
            vals = [float(timestep)] * b
        return (C.c_float * b)(*vals)

    def _check_input(self, x):
        if not x.is_cuda:
            raise RuntimeError("eccv2022-rife_b200: CUDA tensors required -- the B200 hot path has no CPU fallback")
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16, torch.uint8):
            raise TypeError("eccv2022-rife_b200: float32 / float16 / bfloat16 NCHW frames or uint8 HWC frames required")

    @staticmethod
    def _io(dtype):
        return {torch.float32: _lib.RIFE_F32, torch.float16: _lib.RIFE_F16, torch.uint8: _lib.RIFE_U8}[dtype]

    @staticmethod
    def padded_size(h, w, scale):
        """The network needs H, W multiples of 64/scale; the reference's callers zero-pad right/bottom and
        crop the result (inference_img.py:62-67,111).  Here the first kernel pads and the last one crops."""
        if isinstance(scale, (tuple, list)):
            m = int(max(4.0, 4.0 * max(scale)))
        else:
            m = int(round(64 / float(scale)))
        return (h + m - 1) // m * m, (w + m - 1) // m * m

    def plan_for(self, device, prec, b, c, h, w, scale, slot=0, pad=True, nt=1, seq=False):
        hp, wp = self.padded_size(h, w, scale) if pad else (h, w)
        pl = self._get_plan(device, prec, b, c, hp, wp, scale, slot, nt, seq)
        if (hp, wp) != (h, w) or getattr(pl, "frame", (hp, wp)) != (h, w):
            _lib.check(_lib.load().rife_plan_set_frame_size(pl.handle, h, w))
        pl.frame = (h, w)
        return pl

    def run(self, img0, img1, timestep, scale, want_all, pad=False):
        """Shared engine call.  Returns merged[-1] (want_all=False) or the full forward structure.
        float tensors are NCHW (B,C,H,W); uint8 tensors are HWC (B,H,W,C) frames (divided by 255 on the way in,
        `(x*255).byte()` on the way out).  pad=True accepts any H, W (zero-pad + crop inside the kernels)."""
        self._check_input(img0)
        lib = _lib.load()
        if img0.dtype == torch.bfloat16 and torch.is_tensor(img1) and img1.dtype == torch.bfloat16:
            prec0 = self._resolve_precision(torch.bfloat16)
            if prec0 == _lib.PRECISION_FP32:
                a, b_ = img0.float(), img1.float()
            else:
                a, b_ = img0.half(), img1.half()
            r = self._run_prec(a, b_, timestep, scale, want_all, pad, prec0)
            if not want_all:
                return r.to(torch.bfloat16)
            flows, mask, merged = r
            bf = torch.bfloat16
            return ([f.to(bf) for f in flows], mask.to(bf),
                    [(m[0].to(bf), m[1].to(bf)) for m in merged[:4]] + [merged[4].to(bf)])
        return self._run_prec(img0, img1, timestep, scale, want_all, pad, None)

    def _run_prec(self, img0, img1, timestep, scale, want_all, pad, prec):
        lib = _lib.load()
        u8 = img0.dtype == torch.uint8
        if u8:
            b, h, w, c = img0.shape
        else:
            b, c, h, w = img0.shape
        if not torch.is_tensor(img1) or img1.shape != img0.shape or img1.dtype != img0.dtype or img1.device != img0.device:
            raise ValueError("eccv2022-rife_b200: img0 and img1 must have the same shape, dtype and CUDA device "
                             f"(got {tuple(img0.shape)} {img0.dtype} {img0.device} vs "
                             f"{tuple(getattr(img1, 'shape', ()))} {getattr(img1, 'dtype', None)} {getattr(img1, 'device', None)})")
        img0 = img0.contiguous()
        img1 = img1.contiguous()
        dev = img0.device
        if prec is None:
            prec = self._resolve_precision(torch.float16 if u8 else img0.dtype)
        pl = self.plan_for(dev, prec, b, c, h, w, scale, 0, pad or u8)
        io = self._io(img0.dtype)
        ts = self._timesteps(timestep, b)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            if not want_all:
                out = torch.empty_like(img0)
                _lib.check(lib.rife_inference(pl.handle, img0.data_ptr(), img1.data_ptr(), ts, out.data_ptr(), io, stream))
                return out
            if u8:
                raise TypeError("IFNet.forward returns float flows / masks: pass float frames")
            o = _lib.IFNetOutputs()
            flows = [torch.empty((b, 4, h, w), dtype=img0.dtype, device=dev) for _ in range(5)]
            w0 = [torch.empty_like(img0) for _ in range(5)]
            w1 = [torch.empty_like(img0) for _ in range(5)]
            mask = torch.empty((b, 1, h, w), dtype=img0.dtype, device=dev)
            merged = torch.empty_like(img0)
            for i in range(5):
                o.flow[i] = flows[i].data_ptr()
                o.warped0[i] = w0[i].data_ptr()
                o.warped1[i] = w1[i].data_ptr()
            o.mask = mask.data_ptr()
            o.merged = merged.data_ptr()
            _lib.check(lib.rife_ifnet_forward(pl.handle, img0.data_ptr(), img1.data_ptr(), ts, C.byref(o), io, stream))
        merged_list = [(w0[i], w1[i]) for i in range(4)] + [merged]
        return flows, mask, merged_list

    def run_multi(self, img0, img1, timesteps, scale=1.0, pad=False):
        """Several timesteps per frame pair with the Head features computed once per frame.
        img0 / img1: P frame pairs -- or img1=None and img0 = one run of P+1 consecutive frames (video order; each
        frame's features are then shared by its two pairs).  `timesteps`: T floats applied to every pair.
        Returns (P, T, C, H, W) float or (P, T, H, W, C) uint8: out[p, t] = inference(pair p, timesteps[t])."""
        self._check_input(img0)
        lib = _lib.load()
        seq = img1 is None
        if img0.dtype == torch.bfloat16:
            p0 = self._resolve_precision(torch.bfloat16)
            cast = torch.float32 if p0 == _lib.PRECISION_FP32 else torch.float16
            saved, self.precision = self.precision, {_lib.PRECISION_FP32: "fp32", _lib.PRECISION_FP16: "fp16",
                                                     _lib.PRECISION_BF16: "bf16"}[p0]
            try:
                return self.run_multi(img0.to(cast), None if seq else img1.to(cast), timesteps, scale, pad).to(torch.bfloat16)
            finally:
                self.precision = saved
        u8 = img0.dtype == torch.uint8
        if u8:
            n, h, w, c = img0.shape
        else:
            n, c, h, w = img0.shape
        pairs = n - 1 if seq else n
        if pairs < 1:
            raise ValueError("sequence mode needs at least two frames")
        if not seq and (img1.shape != img0.shape or img1.dtype != img0.dtype or img1.device != img0.device):
            raise ValueError("img0 and img1 must have the same shape, dtype and device")
        tl = [float(t) for t in timesteps]
        nt = len(tl)
        if nt < 1:
            raise ValueError("at least one timestep")
        img0 = img0.contiguous()
        img1 = None if seq else img1.contiguous()
        dev = img0.device
        prec = self._resolve_precision(torch.float16 if u8 else img0.dtype)
        pl = self.plan_for(dev, prec, pairs, c, h, w, scale, 0, pad or u8, nt, seq)
        io = self._io(img0.dtype)
        ts = (C.c_float * (pairs * nt))(*(tl * pairs))
        shape = (pairs, nt, h, w, c) if u8 else (pairs, nt, c, h, w)
        with torch.cuda.device(dev):
            out = torch.empty(shape, dtype=img0.dtype, device=dev)
            _lib.check(lib.rife_inference(pl.handle, img0.data_ptr(), None if seq else img1.data_ptr(), ts, out.data_ptr(), io,
                                          torch.cuda.current_stream().cuda_stream))
        return out

    def forward(self, x, timestep=0.5, scale_list=[8, 4, 2, 1], training=False, fastmode=True, ensemble=False):
        """Same signature / return structure as IFNet_HDv3.py:119-169 (inference branch):
        (flow_list[5], mask logits of the last block, merged = 4 (warped0, warped1) tuples + blended tensor)."""
        if training:
            raise NotImplementedError("training=True is broken in the reference (img0 undefined) and out of scope")
        if ensemble:
            print("warning: ensemble is not supported since RIFEv4.21")
        if not fastmode:
            print("contextnet is removed")
        channel = x.shape[1] // 2
        scale = self._scale_from_list(scale_list)
        return self.run(x[:, :channel], x[:, channel:], timestep, scale, want_all=True)
