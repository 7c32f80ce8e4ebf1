"""ECCV-paper RIFE network (reference `model/IFNet.py`, `model/refine.py`, `model/RIFE.py`; BASELINE config 5).

Parameter containers with the reference's state-dict keys / construction order, and a forward that issues the whole
network through the generic NHWC layer API of librife_b200.so (`rife_layer_*`, `rife_nhwc_*`, `rife_eccv_merge`):
activations `[B][H][W][pitch]`, every `torch.cat` of the reference replaced by producers writing into channel
slices of the consumer's tensor.  PyTorch only allocates the buffers.  No CPU fallback.

Two precisions, like the HDv3 path: "fp32" (fp32 tensors, FMA convs, <= 1e-4 vs the reference) and "fp16" (fp16 conv
activations / weights on the tcgen05 kernel with fp32 accumulation; the full-resolution flow / mask / warped-frame
tensor and every flow used by a warp stay fp32).  "auto" picks fp16 for fp16 frames.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .ifnet import _precision_from_env
from .warplayer import base_grids


def _conv_prelu(cin, cout, stride=1):
    # model/IFNet.py:13-18 / model/refine.py:11-16: Sequential(Conv2d, PReLU(out))
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1, bias=True), nn.PReLU(cout))


def _deconv_prelu(cin, cout):
    # model/refine.py:18-22
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=True), nn.PReLU(cout))


class IFBlock(nn.Module):
    """Parameters of model/IFNet.py:20-37."""

    def __init__(self, in_planes, c=64, stride0=2):
        super().__init__()
        self.conv0 = nn.Sequential(_conv_prelu(in_planes, c // 2, stride0), _conv_prelu(c // 2, c, 2))
        self.convblock = nn.Sequential(*[_conv_prelu(c, c) for _ in range(8)])
        self.lastconv = nn.ConvTranspose2d(c, 5, 4, 2, 1)


class Conv2(nn.Module):
    """model/refine.py:24-33."""

    def __init__(self, cin, cout, stride=2):
        super().__init__()
        self.conv1 = _conv_prelu(cin, cout, stride)
        self.conv2 = _conv_prelu(cout, cout, 1)


_C = 16


class Contextnet(nn.Module):
    """Parameters of model/refine.py:36-42."""

    def __init__(self, stride1=2):
        super().__init__()
        self.conv1 = Conv2(3, _C, stride1)
        self.conv2 = Conv2(_C, 2 * _C)
        self.conv3 = Conv2(2 * _C, 4 * _C)
        self.conv4 = Conv2(4 * _C, 8 * _C)


class Unet(nn.Module):
    """Parameters of model/refine.py:59-70."""

    def __init__(self, two_r=False):
        super().__init__()
        self.down0 = Conv2(17, 2 * _C, 1 if two_r else 2)
        self.down1 = Conv2(4 * _C, 4 * _C)
        self.down2 = Conv2(8 * _C, 8 * _C)
        self.down3 = Conv2(16 * _C, 16 * _C)
        self.up0 = _deconv_prelu(32 * _C, 8 * _C)
        self.up1 = _deconv_prelu(16 * _C, 4 * _C)
        self.up2 = _deconv_prelu(8 * _C, 2 * _C)
        self.up3 = _deconv_prelu(4 * _C, _C)
        self.conv = nn.Conv2d(_C, 3, 3, 2 if two_r else 1, 1)


def _pad8(c):
    return (c + 7) // 8 * 8


def _padc(c, h16):
    """Stored channel count of an activation tensor: fp32 -> multiple of 8; fp16 -> 16, or a multiple of 32 so that the
    tcgen05 kernel consumes K in 32- / 64-channel TMA boxes (conv_umma.cu)."""
    if not h16:
        return _pad8(c)
    return 16 if c <= 16 else (c + 31) // 32 * 32


def _dt(t):
    return _lib.RIFE_F16 if t.dtype == torch.float16 else _lib.RIFE_F32


class _Layer:
    """One packed conv layer (rife_layer) plus its geometry."""

    def __init__(self, lib, kind, conv, prelu, cin_store, prec=0):
        w = conv.weight.detach().to("cpu", torch.float32).contiguous()
        b = conv.bias.detach().to("cpu", torch.float32).contiguous() if conv.bias is not None else None
        s = prelu.weight.detach().to("cpu", torch.float32).contiguous() if prelu is not None else None
        self.kind = kind
        self.cin = w.shape[1] if kind != 2 else w.shape[0]
        self.cout = w.shape[0] if kind != 2 else w.shape[1]
        self.cin_store = cin_store
        self.h = C.c_void_p()
        self.lib = lib
        _lib.check(lib.rife_layer_create(C.byref(self.h), kind, w.data_ptr(), b.data_ptr() if b is not None else None,
                                         s.data_ptr() if s is not None else None, self.cin, self.cout, cin_store, prec))
        self.esz = 2 if prec == _lib.PRECISION_FP16 else 4

    def run(self, src, out, hh, ww, act, res=None, out_off=0, stream=0):
        """src / out / res: NHWC tensors of the layer's precision; the conv reads channels [0, cin_store) of src and
        writes channels [out_off, ...) of out."""
        b = src.shape[0]
        _lib.check(self.lib.rife_layer_run(self.h, src.data_ptr(), src.shape[-1], out.data_ptr() + out.element_size() * out_off,
                                           _dt(out), out.shape[-1],
                                           b, hh, ww, res.data_ptr() if res is not None else None,
                                           res.shape[-1] if res is not None else 0, act, stream))

    def free(self):
        if self.h:
            self.lib.rife_layer_destroy(self.h)
            self.h = None


class IFNet(nn.Module):
    """Drop-in for `model.IFNet.IFNet` (inference branch: `x` holds img0, img1 and no ground truth)."""

    TIMESTEP_PLANE = 0      # IFNet_m adds one constant timestep channel after img1 (model/IFNet_m.py:56-59,64)
    TWO_R = False           # IFNet_2R: the stride variant of model/IFNet_2R.py + model/refine_2R.py

    def __init__(self, precision="auto"):
        super().__init__()
        self.precision = _precision_from_env(precision)
        tp = self.TIMESTEP_PLANE
        s0 = 1 if self.TWO_R else 2                   # model/IFNet_2R.py:24
        self.block0 = IFBlock(6 + tp, c=240, stride0=s0)   # model/IFNet.py:56-61: same construction order => same default init
        self.block1 = IFBlock(13 + 4 + tp, c=150, stride0=s0)
        self.block2 = IFBlock(13 + 4 + tp, c=90, stride0=s0)
        self.block_tea = IFBlock(16 + 4 + tp, c=90, stride0=s0)   # teacher: parameters only (training branch, model/IFNet.py:88-97)
        self.contextnet = Contextnet(s0)
        self.unet = Unet(self.TWO_R)
        self._layers = {}      # precision -> {name: _Layer}
        self._arenas = {}      # (device, B, H, W, fp16?, scales) -> work tensors of one forward, in call order
        self._stamp = None
        self._generation = 0

    # ---------------------------------------------------------------- packed-layer cache
    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._generation += 1
        return r

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._generation += 1
        return r

    def invalidate(self):
        self._generation += 1

    def release(self):
        for layers in (self._layers or {}).values():
            for l in layers.values():
                l.free()
        self._layers = {}
        self._arenas = {}

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _pack(self, device, prec):
        stamp = (self._generation, sum(p._version for p in self.parameters()), str(device))
        if self._stamp != stamp:
            self.release()
            self._stamp = stamp
        if prec in self._layers:
            return self._layers[prec]
        lib = _lib.load()
        h16 = prec == _lib.PRECISION_FP16
        pc = lambda c: _padc(c, h16)
        mk = lambda kind, conv, prelu, cs: _Layer(lib, kind, conv, prelu, cs, prec)
        L = {}
        with torch.cuda.device(device):
            for i, blk in enumerate((self.block0, self.block1, self.block2)):
                c = blk.lastconv.weight.shape[0]
                cin = blk.conv0[0][0].weight.shape[1]
                L[f"b{i}.c00"] = (mk(0, blk.conv0[0][0], blk.conv0[0][1], pc(cin)) if self.TWO_R
                                  else mk(1, blk.conv0[0][0], blk.conv0[0][1], _pad8(cin)))
                L[f"b{i}.c01"] = mk(1, blk.conv0[1][0], blk.conv0[1][1], pc(c // 2))
                for j in range(8):
                    L[f"b{i}.cb{j}"] = mk(0, blk.convblock[j][0], blk.convblock[j][1], pc(c))
                L[f"b{i}.last"] = mk(2, blk.lastconv, None, pc(c))
            cn = self.contextnet
            for k, (m, cin) in enumerate(((cn.conv1, 3), (cn.conv2, _C), (cn.conv3, 2 * _C), (cn.conv4, 4 * _C))):
                L[f"ctx{k}.1"] = (mk(0, m.conv1[0], m.conv1[1], pc(cin)) if (self.TWO_R and k == 0)
                                  else mk(1, m.conv1[0], m.conv1[1], _pad8(cin)))
                L[f"ctx{k}.2"] = mk(0, m.conv2[0], m.conv2[1], pc(m.conv2[0].weight.shape[1]))
            un = self.unet
            for k, (m, cin) in enumerate(((un.down0, 17), (un.down1, 4 * _C), (un.down2, 8 * _C), (un.down3, 16 * _C))):
                L[f"dn{k}.1"] = (mk(0, m.conv1[0], m.conv1[1], pc(cin)) if (self.TWO_R and k == 0)
                                  else mk(1, m.conv1[0], m.conv1[1], _pad8(cin)))
                L[f"dn{k}.2"] = mk(0, m.conv2[0], m.conv2[1], pc(m.conv2[0].weight.shape[1]))
            for k, m in enumerate((un.up0, un.up1, un.up2, un.up3)):
                L[f"up{k}"] = mk(2, m[0], m[1], m[0].weight.shape[0])
            L["out"] = mk(1 if self.TWO_R else 0, un.conv, None, _C)      # refine_2R.py:71: stride 2 (from 2x resolution)
        self._layers[prec] = L
        return L

    # ---------------------------------------------------------------- forward
    def forward(self, x, scale=[4, 2, 1], timestep=0.5, returnflow=False):
        """model/IFNet.py:63-108, inference branch.  Returns the reference's 6-tuple
        (flow_list[3], sigmoid(mask), merged[3], None, None, 0).  `timestep` is ignored by this network; the IFNet_m
        subclass feeds it to every block as a constant plane (float, or a (B,1,1,1) tensor) and honours `returnflow`
        (model/IFNet_m.py:63-64,104-105)."""
        return self._run(x, None, scale, timestep, returnflow)

    def _run(self, x, img1, scale=[4, 2, 1], timestep=0.5, returnflow=False):
        """`img1` None: `x` is the reference's (B, 6, H, W) cat(img0, img1); else `x` = img0 and `img1` two (B, 3, H, W)
        frames, packed straight into the working tensor (the `torch.cat` of model/RIFE.py:59 never happens)."""
        if img1 is not None:
            if img1.shape != x.shape or img1.dtype != x.dtype or img1.device != x.device or x.shape[1] != 3:
                raise ValueError("img0 and img1 must be (B, 3, H, W) tensors of one dtype on one device")
            img1 = img1.contiguous()
        if not x.is_cuda:
            raise RuntimeError("eccv2022-rife_b200: CUDA tensors required -- the B200 hot path has no CPU fallback")
        if x.dtype not in (torch.float32, torch.float16):
            raise TypeError("float32 or float16 frames required")
        if img1 is None and x.shape[1] != 6:
            raise NotImplementedError("training branch (ground truth appended to x, model/IFNet.py:66,88-97) is out of scope")
        s3 = [float(v) for v in scale[:3]]
        if any(v not in (1.0, 2.0, 4.0, 8.0, 16.0) for v in s3):
            raise NotImplementedError(f"scale entries must be powers of two >= 1, got {list(scale)}")
        B, _, H, W = x.shape
        tp = self.TIMESTEP_PLANE
        if tp:
            if torch.is_tensor(timestep):
                tvals = [float(v) for v in timestep.detach().reshape(-1).cpu().tolist()]
                tvals = tvals * B if len(tvals) == 1 else tvals
                if len(tvals) != B:
                    raise ValueError("tensor timestep must broadcast as (B,1,1,1)")
            else:
                tvals = [float(timestep)] * B
        two_r = self.TWO_R
        need = int(max(8, 2 * max(s3))) if two_r else int(max(16, 4 * max(s3)))
        if H % need or W % need:
            raise RuntimeError(f"ECCV IFNet: H and W must be multiples of {need} for scale list {list(scale)} (got {H}x{W})")
        lib = _lib.load()
        dev = x.device
        mode = self.precision
        if mode == "auto":
            mode = "fp16" if x.dtype == torch.float16 else "fp32"
        h16 = mode == "fp16"
        L = self._pack(dev, _lib.PRECISION_FP16 if h16 else _lib.PRECISION_FP32)
        pc = lambda c: _padc(c, h16)
        adt = torch.float16 if h16 else torch.float32       # conv activations
        x = x.contiguous()
        io = _lib.RIFE_F32 if x.dtype == torch.float32 else _lib.RIFE_F16
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            # work tensors come from a per-(shape, precision) arena in call order: allocated and zero-filled once, then
            # re-used by every later forward (zero padding channels are never written, so they stay zero)
            akey = (str(dev), B, H, W, h16, tuple(s3))
            if akey not in self._arenas and len(self._arenas) >= 4:
                self._arenas.pop(next(iter(self._arenas)))        # keep the work memory of a few recent shapes only
            arena = self._arenas.setdefault(akey, [])
            cursor = [0]

            def z(h, w, p, dt=adt):
                k = cursor[0]
                cursor[0] += 1
                if k == len(arena):
                    arena.append(torch.zeros((B, h, w, p), dtype=dt, device=dev))
                t = arena[k]
                assert t.shape == (B, h, w, p) and t.dtype == dt
                return t

            e = z
            f32 = torch.float32
            chk = _lib.check

            def resize(src, soff, c, dst, doff, mul=1.0, acc=0):
                chk(lib.rife_nhwc_resize(src.data_ptr(), _dt(src), src.shape[-1], soff, dst.data_ptr(), _dt(dst), dst.shape[-1],
                                         doff, B, c, src.shape[1], src.shape[2], dst.shape[1], dst.shape[2], mul, acc, st))

            def warp(src, soff, c, fl, foff, dst, doff):
                h, w = src.shape[1], src.shape[2]
                _, _, gx, gy = base_grids(h, w, dev)
                chk(lib.rife_nhwc_warp(src.data_ptr(), _dt(src), src.shape[-1], soff, fl.data_ptr(), _dt(fl), fl.shape[-1], foff,
                                       dst.data_ptr(), _dt(dst), dst.shape[-1], doff, B, c, h, w, gx.data_ptr(), gy.data_ptr(), st))

            def s2d(src, c, cpad, soff=0):
                h, w = src.shape[1], src.shape[2]
                dst = e(h // 2, w // 2, 4 * cpad)
                chk(lib.rife_nhwc_s2d(src.data_ptr(), _dt(src), src.shape[-1], soff, c, dst.data_ptr(), _dt(dst), cpad, B, h, w, st))
                return dst

            def conv2(prefix, src, c_in, dst, doff, soff=0, stride1=False):
                """Conv2 (model/refine.py:24-33): stride-2 conv + PReLU, conv + PReLU; result into dst[..., doff:].
                stride1: the first conv of refine_2R.py's Contextnet.conv1 / Unet.down0."""
                l1, l2 = L[prefix + ".1"], L[prefix + ".2"]
                if stride1:
                    h, w = src.shape[1], src.shape[2]
                    xi = z(h, w, l1.cin_store)
                    resize(src, soff, c_in, xi, 0)                                    # same size: a copy into the conv's
                    a = z(h, w, l2.cin_store)                                         # own (padded, activation-dtype) tensor
                    l1.run(xi, a, h, w, 2, stream=st)
                    l2.run(a, dst, h, w, 2, out_off=doff, stream=st)
                    return
                h, w = src.shape[1] // 2, src.shape[2] // 2
                a = z(h, w, l2.cin_store)
                l1.run(s2d(src, c_in, l1.cin_store, soff), a, h, w, 2, stream=st)
                l2.run(a, dst, h, w, 2, out_off=doff, stream=st)

            # full-resolution fp32 working tensor in Unet.down0's cat order (model/refine.py:73):
            # [img0 0:3 | img1 3:6 | warped0 6:9 | warped1 9:12 | mask 12 | flow 13:17 | pad]
            full = z(H, W, 24, f32)
            if img1 is None:
                chk(lib.rife_nhwc_from_nchw(x.data_ptr(), io, full.data_ptr(), _dt(full), B, 6, H, W, 24, 0, st))
            else:
                chk(lib.rife_nhwc_from_nchw(x.data_ptr(), io, full.data_ptr(), _dt(full), B, 3, H, W, 24, 0, st))
                chk(lib.rife_nhwc_from_nchw(img1.data_ptr(), io, full.data_ptr(), _dt(full), B, 3, H, W, 24, 3, st))
            flow_list, merged, mask_sig = [], [], None
            out_dt = x.dtype
            for i in range(3):                                                        # model/IFNet.py:74-86
                s = int(s3[i])
                hs, ws = H // s, W // s
                cin = (6 if i == 0 else 17) + tp
                xin = z(hs, ws, pc(cin) if two_r else _pad8(cin))
                if tp:                                                                # IFNet_m.py:78,82: [img0,img1,t,...]
                    resize(full, 0, 6, xin, 0)
                    if len(set(tvals)) == 1:
                        chk(lib.rife_nhwc_fill(xin.data_ptr(), _dt(xin), xin.shape[-1], 6, B, 1, hs, ws, tvals[0], st))
                    else:
                        for bi, tv in enumerate(tvals):
                            chk(lib.rife_nhwc_fill(xin[bi].data_ptr(), _dt(xin), xin.shape[-1], 6, 1, 1, hs, ws, tv, st))
                    if i > 0:
                        resize(full, 6, 7, xin, 7)                                    # warped0, warped1, mask
                        resize(full, 13, 4, xin, 14, mul=1.0 / s)
                elif i == 0:
                    resize(full, 0, 6, xin, 0)                                        # :40 (identity when scale == 1)
                else:
                    resize(full, 0, 13, xin, 0)                                       # cat(img0,img1,warped0,warped1,mask)
                    resize(full, 13, 4, xin, 13, mul=1.0 / s)                         # :42 interp(flow) * 1/scale
                c = (240, 150, 90)[i]
                if two_r:                                                             # IFNet_2R.py:24: conv0[0] at stride 1
                    a = z(hs, ws, pc(c // 2))
                    L[f"b{i}.c00"].run(xin, a, hs, ws, 2, stream=st)
                    hb, wb = hs // 2, ws // 2
                else:
                    a = z(hs // 2, ws // 2, pc(c // 2))
                    L[f"b{i}.c00"].run(s2d(xin, cin, _pad8(cin)), a, hs // 2, ws // 2, 2, stream=st)
                    hb, wb = hs // 4, ws // 4
                bq = z(hb, wb, pc(c))
                L[f"b{i}.c01"].run(s2d(a, pc(c // 2), pc(c // 2)), bq, hb, wb, 2, stream=st)
                pp = (z(hb, wb, pc(c)), z(hb, wb, pc(c)))
                y = bq
                for j in range(8):                                                    # :45-46 convblock(x) + x
                    dst = pp[j & 1]
                    L[f"b{i}.cb{j}"].run(y, dst, hb, wb, 3 if j == 7 else 2, res=bq if j == 7 else None, stream=st)
                    y = dst
                tmp = e(2 * hb, 2 * wb, 8, f32)                                       # flow / mask head: fp32 in both modes
                L[f"b{i}.last"].run(y, tmp, hb, wb, 0, stream=st)                     # :47 lastconv -> 5 ch at 1/(2s) (2R: 1/s)
                up_mul = float(s if two_r else 2 * s)                                 # :48-49 (IFNet_2R.py:48-49: scale only)
                resize(tmp, 0, 4, full, 13, mul=up_mul, acc=int(i > 0))               # :76: flow (+)= up(tmp) * up factor
                resize(tmp, 4, 1, full, 12, mul=1.0, acc=int(i > 0))                  # :50,77: mask (+)= up(tmp)
                warp(full, 0, 3, full, 13, full, 6)                                   # :82
                warp(full, 3, 3, full, 15, full, 9)                                   # :83
                fl = torch.empty((B, 4, H, W), dtype=out_dt, device=dev)
                chk(lib.rife_nhwc_to_nchw(full.data_ptr(), _dt(full), 24, 13, fl.data_ptr(), io, B, 4, H, W, 0, st))
                flow_list.append(fl)
                if i < 2:                                                             # :98-99
                    mg = torch.empty((B, 3, H, W), dtype=out_dt, device=dev)
                    chk(lib.rife_eccv_merge(full.data_ptr(), 24, 6, full.data_ptr(), 24, 9, full.data_ptr(), 24, 12, _dt(full),
                                            None, 0, 0, 0, mg.data_ptr(), io, B, H, W, st))
                    merged.append(mg)
            if tp and returnflow:
                return flow_list[2]                                                   # IFNet_m.py:104-105
            mask_sig = torch.empty((B, 1, H, W), dtype=out_dt, device=dev)
            chk(lib.rife_nhwc_to_nchw(full.data_ptr(), _dt(full), 24, 12, mask_sig.data_ptr(), io, B, 1, H, W, 1, st))

            # Contextnet on both frames (model/refine.py:44-57), features warped straight into the Unet's cat slices
            q = 1 if two_r else 2                         # refine_2R.py: every level of the refinement one octave finer
            d1, d2, d3, u0 = (z(H // q, W // q, 64), z(H // (2 * q), W // (2 * q), 128), z(H // (4 * q), W // (4 * q), 256),
                              z(H // (8 * q), W // (8 * q), 512))
            for f, (img_off, flow_off) in enumerate(((0, 13), (3, 15))):
                if two_r:                                                             # refine_2R.py:45-48: full-res flow
                    fl, foff = full, flow_off
                else:
                    fl, foff = z(H // 2, W // 2, 8, f32), 0
                    resize(full, flow_off, 2, fl, 0, mul=0.5)
                feat = e(H // q, W // q, 16)
                conv2("ctx0", full, 3, feat, 0, soff=img_off, stride1=two_r)
                warp(feat, 0, 16, fl, foff, d1, 32 + 16 * f)
                prev, prev_c = feat, 16
                for k, (dst, base) in enumerate(((d2, 64), (d3, 128), (u0, 256))):
                    cc = 32 << k
                    h, w = prev.shape[1] // 2, prev.shape[2] // 2
                    nxt = e(h, w, cc)
                    conv2(f"ctx{k + 1}", prev, prev_c, nxt, 0)
                    fl2 = z(h, w, 8, f32)
                    resize(fl, foff, 2, fl2, 0, mul=0.5)
                    warp(nxt, 0, cc, fl2, 0, dst, base + cc * f)
                    prev, prev_c, fl, foff = nxt, cc, fl2, 0

            # Unet (model/refine.py:72-82)
            conv2("dn0", full, 17, d1, 0, stride1=two_r)
            conv2("dn1", d1, 64, d2, 0)
            conv2("dn2", d2, 128, d3, 0)
            conv2("dn3", d3, 256, u0, 0)
            u1, u2, u3 = e(H // (4 * q), W // (4 * q), 256), e(H // (2 * q), W // (2 * q), 128), e(H // q, W // q, 64)
            L["up0"].run(u0, u1, H // (8 * q), W // (8 * q), 2, stream=st)
            resize(d3, 0, 128, u1, 128)                                               # cat((x, s2))
            L["up1"].run(u1, u2, H // (4 * q), W // (4 * q), 2, stream=st)
            resize(d2, 0, 64, u2, 64)
            L["up2"].run(u2, u3, H // (2 * q), W // (2 * q), 2, stream=st)
            resize(d1, 0, 32, u3, 32)
            xo = e(2 * H // q, 2 * W // q, 16)
            L["up3"].run(u3, xo, H // q, W // q, 2, stream=st)
            ref = e(H, W, pc(3), f32)
            if two_r:
                L["out"].run(s2d(xo, 16, 16), ref, H, W, 0, stream=st)                # refine_2R.py:71: stride 2 back to (H, W)
            else:
                L["out"].run(xo, ref, H, W, 0, stream=st)
            mg = torch.empty((B, 3, H, W), dtype=out_dt, device=dev)
            chk(lib.rife_eccv_merge(full.data_ptr(), 24, 6, full.data_ptr(), 24, 9, full.data_ptr(), 24, 12, _dt(full),
                                    ref.data_ptr(), _dt(ref), ref.shape[-1], 0, mg.data_ptr(), io, B, H, W, st))
            merged.append(mg)                                                          # :105-107
        return flow_list, mask_sig, merged, None, None, 0


class IFNet_m(IFNet):
    """Drop-in for `model.IFNet_m.IFNet_m`: the arbitrary-timestep variant (benchmark/HD_multi_4X.py:16)."""
    TIMESTEP_PLANE = 1


class IFNet_2R(IFNet):
    """Drop-in for `model.IFNet_2R.IFNet` (with `model.refine_2R`): IFBlock.conv0[0], Contextnet.conv1 and Unet.down0 at
    stride 1, the Unet's output conv at stride 2 -- every block and the refinement work one octave finer.  Same 190
    tensors and default init as IFNet; no Model class of the reference selects it, so it is used through forward()."""
    TWO_R = True


device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class Model:
    """Drop-in for `model.RIFE.Model` (inference side, model/RIFE.py:18-65)."""

    def __init__(self, local_rank=-1, arbitrary=False, precision="auto"):
        if local_rank != -1:
            raise NotImplementedError("DDP training wrapper is out of scope: inference only")
        self.flownet = IFNet_m(precision=precision) if arbitrary else IFNet(precision=precision)   # model/RIFE.py:20-23
        self.device()

    def train(self):
        self.flownet.train()

    def eval(self):
        self.flownet.eval()

    def device(self):
        self.flownet.to(device)

    def load_model(self, path, rank=0):
        """model/RIFE.py:41-50: keeps only 'module.'-prefixed keys, strips the prefix, strict load."""
        if rank <= 0:
            sd = torch.load("{}/flownet.pkl".format(path), map_location="cpu" if not torch.cuda.is_available() else None)
            self.flownet.load_state_dict({k.replace("module.", ""): v for k, v in sd.items() if "module." in k})

    def save_model(self, path, rank=0):
        if rank == 0:
            torch.save(self.flownet.state_dict(), "{}/flownet.pkl".format(path))

    def inference(self, img0, img1, scale=1, scale_list=[4, 2, 1], TTA=False, timestep=0.5):
        """model/RIFE.py:56-65.  (The reference rescales its default `scale_list` argument in place; here a copy is used.)"""
        sl = [scale_list[i] * 1.0 / scale for i in range(3)]
        merged = self.flownet._run(img0, img1, sl, timestep)[2]      # both frames packed by the library: no torch.cat
        if not TTA:
            return merged[2]
        # test-time augmentation (model/RIFE.py:61-65): the same pair flipped in both axes, results averaged
        merged2 = self.flownet._run(img0.flip(2).flip(3), img1.flip(2).flip(3), sl, timestep)[2]
        return (merged[2] + merged2[2].flip(2).flip(3)) / 2
