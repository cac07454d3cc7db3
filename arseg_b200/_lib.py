"""ctypes binding of libarseg_sm100a.so (the C ABI declared in include/arseg.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libarseg_sm100a.so")

OK, E_BADARG, E_UNSUPPORTED, E_CUDA = 0, -1, -2, -3
F32, BF16, F64, I16, F16 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_PRELU = 0, 1, 2
RESIZE_BILINEAR, RESIZE_BILINEAR_AC, RESIZE_NEAREST = 0, 1, 2
NCHW, NHWC = 0, 1
CONV_SIMT_F32, CONV_TC_TF32, CONV_TC_BF16, CONV_TC_F16 = 1, 2, 3, 4
CREFF_EXACT_F32, CREFF_MMA_F16, CREFF_TCGEN05 = 0, 1, 2
CREFF_PHASE_ALL, CREFF_PHASE_PREPASS, CREFF_PHASE_MAIN = 0, 1, 2
ABI_VERSION = 6

vp, ci, cf = C.c_void_p, C.c_int, C.c_float


class ConvDesc(C.Structure):
    _fields_ = [("inp", vp), ("w", vp), ("scale", vp), ("shift", vp), ("residual", vp), ("out", vp),
                ("dtype", ci), ("N", ci), ("Hi", ci), ("Wi", ci), ("Cin", ci), ("Cout", ci), ("KH", ci), ("KW", ci),
                ("stride", ci), ("pad", ci), ("dil", ci), ("out_cstride", ci), ("out_coff", ci), ("act", ci),
                ("prelu_slope", cf), ("engine", ci), ("out_f32", ci)]


class CreffArgs(C.Structure):
    _fields_ = [("hr", vp), ("hr_shared", ci), ("hr_layout", ci), ("engine", ci), ("flow", vp), ("flow_dtype", ci), ("Hm", ci), ("Wm", ci),
                ("lr", vp), ("lr_layout", ci), ("lr_dtype", ci), ("h", ci), ("w", ci),
                ("wq", vp), ("bq", vp), ("wk", vp), ("bk", vp), ("wv", vp), ("bv", vp), ("wcls", vp), ("bcls", vp),
                ("ncls", ci), ("log_softmax", ci), ("out_p", vp), ("out_logits", vp), ("out_argmax", vp),
                ("N", ci), ("C", ci), ("H", ci), ("W", ci), ("k", ci), ("workspace", vp), ("workspace_bytes", C.c_size_t), ("hr_dtype", ci), ("phase", ci)]


_PROTOS = {
    "arseg_abi_version": ([], ci),
    "arseg_last_error": ([], C.c_char_p),
    "arseg_local_similar_fwd": ([vp, vp, vp] + [ci] * 6 + [vp], ci),
    "arseg_local_weighting_fwd": ([vp, vp, vp] + [ci] * 6 + [vp], ci),
    "arseg_local_similar_bwd": ([vp, vp, vp] + [ci] * 7 + [vp], ci),
    "arseg_local_weighting_bwd_ori": ([vp, vp, vp] + [ci] * 6 + [vp], ci),
    "arseg_local_weighting_bwd_weight": ([vp, vp, vp] + [ci] * 6 + [vp], ci),
    "arseg_warp_feature_nchw": ([vp, vp, ci, vp, ci, ci, ci, ci, vp], ci),
    "arseg_resize_nchw_f32": ([vp, vp] + [ci] * 6 + [vp], ci),
    "arseg_resize_argmax_nchw": ([vp, vp, vp] + [ci] * 7 + [vp], ci),
    "arseg_log_softmax_nchw": ([vp, vp, ci, ci, ci, ci, vp], ci),
    "arseg_confusion_hist":([vp, vp, vp, C.c_longlong, ci, ci, vp], ci),
    "arseg_frame_ingest_u8": ([vp, C.POINTER(cf), C.POINTER(cf), vp, ci, ci, ci, ci, ci, ci, vp], ci),
    "arseg_merge_motion_workspace_bytes": ([ci, ci, ci], C.c_size_t),
    "arseg_merge_motion": ([vp, vp, C.c_size_t, vp, ci, ci, ci, vp], ci),
    "arseg_nchw_to_nhwc": ([vp, vp, ci, ci, ci, ci, ci, vp], ci),
    "arseg_nhwc_to_nchw": ([vp, ci, vp, ci, ci, ci, ci, vp], ci),
    "arseg_conv_stem7x7s2": ([vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp], ci),
    "arseg_maxpool3x3s2_nhwc": ([vp, vp, ci, ci, ci, ci, ci, vp], ci),
    "arseg_conv2d_nhwc": ([C.POINTER(ConvDesc), vp], ci),
    "arseg_resize_nhwc": ([vp, vp] + [ci] * 10 + [vp], ci),
    "arseg_adaptive_avgpool_nhwc": ([vp, vp] + [ci] * 8 + [vp], ci),
    "arseg_global_maxpool_nhwc": ([vp, vp] + [ci] * 5 + [vp], ci),
    "arseg_linear_f32": ([vp, vp, vp, vp, ci, ci, ci, ci, vp], ci),
    "arseg_gate_nhwc": ([vp, vp, vp, vp, ci, vp, vp, vp, ci, ci, ci, ci, ci, vp], ci),
    "arseg_pyramid_pool_nhwc": ([vp, vp, ci, ci, ci, ci, ci, vp, ci, vp], ci),
    "arseg_pyramid_conv1x1": ([vp, vp, vp, vp, ci, vp, ci, ci, ci, vp, ci, vp], ci),
    "arseg_pyramid_upsample_concat": ([vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, ci, vp], ci),
    "arseg_creff_fused_fwd": ([C.POINTER(CreffArgs), vp], ci),
    "arseg_creff_workspace_bytes": ([C.POINTER(CreffArgs)], C.c_size_t),
}

EXPORTED_SYMBOLS = tuple(_PROTOS.keys())

_lib = None
_lock = threading.Lock()


class ArsegError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the C-ABI library (built by `python -m arseg_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ArsegError("libarseg_sm100a.so not found at %s -- build it with `python -m arseg_b200.build`; "
                                 "there is no CPU or PyTorch fallback" % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (args, res) in _PROTOS.items():
                fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
                fn.argtypes = args
                fn.restype = res
            if lib.arseg_abi_version() != ABI_VERSION:
                raise ArsegError("libarseg_sm100a.so ABI version mismatch")
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        msg = load().arseg_last_error().decode(errors="replace")
        kind = {E_BADARG: "bad argument", E_UNSUPPORTED: "unsupported", E_CUDA: "CUDA error"}.get(rc, "error %d" % rc)
        raise ArsegError("%s: %s: %s" % (what or "arseg", kind, msg))
