"""ctypes binding of csrc/librife_b200.so (include/rife_b200.h).  Fails loudly when the library is missing:
there is no CPU / PyTorch fallback for the hot path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RIFE_B200_LIB") or os.path.join(_HERE, "csrc", "librife_b200.so")   # override: A/B builds

RIFE_F32, RIFE_F16, RIFE_U8 = 0, 1, 2
PRECISION_FP32, PRECISION_FP16, PRECISION_BF16 = 0, 1, 2

# every symbol declared in include/rife_b200.h
SYMBOLS = [
    "rife_abi_version", "rife_last_error", "rife_launch_count", "rife_weights_create", "rife_weights_destroy",
    "rife_plan_workspace_bytes", "rife_plan_create", "rife_plan_workspace_bytes_ex", "rife_plan_create_ex",
    "rife_plan_workspace_bytes_list", "rife_plan_create_list", "rife_plan_destroy", "rife_plan_set_frame_size",
    "rife_plan_set_profiling", "rife_plan_set_chunk",
    "rife_plan_get_profile", "rife_inference", "rife_ifnet_forward",
    "rife_inference_host", "rife_inference_host_pipelined", "rife_warp", "rife_resize_bilinear", "rife_conv3x3", "rife_conv_transpose4x4s2",
    "rife_scene_ssim", "rife_layer_create", "rife_layer_destroy", "rife_layer_run", "rife_nhwc_from_nchw", "rife_nhwc_to_nchw",
    "rife_nhwc_resize", "rife_nhwc_warp", "rife_nhwc_s2d", "rife_nhwc_fill", "rife_eccv_merge",
]


class RifeError(RuntimeError):
    pass


class IFNetOutputs(C.Structure):
    """struct rife_ifnet_outputs"""
    _fields_ = [("flow", C.c_void_p * 5), ("mask", C.c_void_p), ("warped0", C.c_void_p * 5),
                ("warped1", C.c_void_p * 5), ("merged", C.c_void_p)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RifeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C {os.path.dirname(LIB_PATH)}`.  The RIFE B200 hot path has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    fp = C.POINTER(C.c_float)
    lib.rife_abi_version.restype = i
    lib.rife_last_error.restype = C.c_char_p
    lib.rife_launch_count.restype = C.c_uint64
    lib.rife_weights_create.argtypes = [C.POINTER(vp), i, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(C.c_int64), i]
    lib.rife_weights_destroy.argtypes = [vp]
    lib.rife_plan_workspace_bytes.argtypes = [i, i, i, i, f, i]
    lib.rife_plan_workspace_bytes.restype = sz
    lib.rife_plan_create.argtypes = [C.POINTER(vp), vp, i, i, i, i, f, vp, sz, vp, vp]
    lib.rife_plan_workspace_bytes_ex.argtypes = [i, i, i, i, i, i, f, i]
    lib.rife_plan_workspace_bytes_ex.restype = sz
    lib.rife_plan_create_ex.argtypes = [C.POINTER(vp), vp, i, i, i, i, i, i, f, vp, sz, vp, vp]
    lib.rife_plan_workspace_bytes_list.argtypes = [i, i, i, i, i, i, fp, i]
    lib.rife_plan_workspace_bytes_list.restype = sz
    lib.rife_plan_create_list.argtypes = [C.POINTER(vp), vp, i, i, i, i, i, i, fp, vp, sz, vp, vp]
    lib.rife_plan_destroy.argtypes = [vp]
    lib.rife_plan_set_frame_size.argtypes = [vp, i, i]
    lib.rife_plan_set_profiling.argtypes = [vp, i]
    lib.rife_plan_set_chunk.argtypes = [vp, i]
    lib.rife_plan_get_profile.argtypes = [vp, C.POINTER(C.c_char_p), fp, i]
    lib.rife_inference.argtypes = [vp, vp, vp, fp, vp, i, vp]
    lib.rife_ifnet_forward.argtypes = [vp, vp, vp, fp, C.POINTER(IFNetOutputs), i, vp]
    lib.rife_inference_host.argtypes = [vp, vp, vp, fp, vp, i, vp]
    lib.rife_inference_host_pipelined.argtypes = [vp, vp, vp, fp, vp, i, vp, vp, vp]
    lib.rife_warp.argtypes = [vp, vp, vp, i, i, i, i, i, vp, vp, vp]
    lib.rife_resize_bilinear.argtypes = [vp, vp, i, i, i, i, i, i, f, vp]
    lib.rife_conv3x3.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, vp]
    lib.rife_conv_transpose4x4s2.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, vp]
    lib.rife_layer_create.argtypes = [C.POINTER(vp), i, vp, vp, vp, i, i, i, i]
    lib.rife_layer_destroy.argtypes = [vp]
    lib.rife_layer_run.argtypes = [vp, vp, i, vp, i, i, i, i, i, vp, i, i, vp]
    lib.rife_nhwc_from_nchw.argtypes = [vp, i, vp, i, i, i, i, i, i, i, vp]
    lib.rife_nhwc_to_nchw.argtypes = [vp, i, i, i, vp, i, i, i, i, i, i, vp]
    lib.rife_nhwc_resize.argtypes = [vp, i, i, i, vp, i, i, i, i, i, i, i, i, i, f, i, vp]
    lib.rife_nhwc_warp.argtypes = [vp, i, i, i, vp, i, i, i, vp, i, i, i, i, i, i, i, vp, vp, vp]
    lib.rife_nhwc_s2d.argtypes = [vp, i, i, i, i, vp, i, i, i, i, i, vp]
    lib.rife_scene_ssim.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp]
    lib.rife_nhwc_fill.argtypes = [vp, i, i, i, i, i, i, i, f, vp]
    lib.rife_eccv_merge.argtypes = [vp, i, i, vp, i, i, vp, i, i, i, vp, i, i, i, vp, i, i, i, i, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("rife_abi_version",):
            pass
    if lib.rife_abi_version() != 1:
        raise RifeError("librife_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().rife_last_error()
        raise RifeError(f"librife_b200 error {rc}: {msg.decode() if msg else '?'}")
