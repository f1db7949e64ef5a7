"""ctypes binding of the C ABI declared in include/maxstyle_b200.h.

There is deliberately no CPU or eager-PyTorch fallback: if the compiled library is missing the
compute entry points raise, loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libmaxstyle_b200.so")

# enums of include/maxstyle_b200.h
OK = 0
ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA, ERR_NO_DEVICE, ERR_TIMEOUT = 1, 2, 3, 4, 5, 6
F32, BF16 = 0, 1
NCHW, NHWC = 0, 1
FLAG_MIX_STYLE, FLAG_NO_NOISE, FLAG_COMPUTE_BATCH_STD, FLAG_NO_CLAMP = 1, 2, 4, 8
STEP_NONE, STEP_ADAM, STEP_SIGN = 0, 1, 2
PRE_NONE, PRE_LEAKY_RELU, PRE_SIGMOID = 0, 1, 2
SWEEP_REVERSE, SWEEP_X_KEEP, SWEEP_X_STREAM, SWEEP_IO_NORMAL, SWEEP_NO_FUSED = 1, 2, 4, 8, 16
SWEEP_NO_RESIDENT, SWEEP_FORCE_WINDOW, SWEEP_FORCE_RESIDENT, SWEEP_NO_RING, SWEEP_FORCE_RING = 32, 64, 128, 256, 512
SWEEP_NO_CLUSTER, SWEEP_FORCE_CLUSTER = 1024, 2048
SWEEP_CLUSTER_SIZE_SHIFT, SWEEP_CLUSTER_STAGES_SHIFT, SWEEP_CLUSTER_PIECES_SHIFT = 12, 16, 19
SWEEP_NO_PAIR, SWEEP_FORCE_PAIR = 1 << 25, 1 << 26

_f32p = C.c_void_p      # device pointers travel as plain addresses
_vp = C.c_void_p


class StepStruct(C.Structure):
    """maxstyle_step_t"""
    _fields_ = [
        ("mode", C.c_int32), ("maximize", C.c_int32),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("t", C.c_int32), ("update_noise", C.c_int32), ("update_mix", C.c_int32), ("reserved", C.c_int32),
        ("step_dev", _vp),
        ("gamma_noise", _f32p), ("beta_noise", _f32p), ("lmda", _f32p),
        ("gamma_m", _f32p), ("gamma_v", _f32p), ("beta_m", _f32p), ("beta_v", _f32p),
        ("lmda_m", _f32p), ("lmda_v", _f32p),
    ]


# name -> (restype, argtypes); every symbol the header declares
SIGNATURES = {
    "maxstyle_version": (C.c_char_p, []),
    "maxstyle_strerror": (C.c_char_p, [C.c_int]),
    "maxstyle_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "maxstyle_stats": (C.c_int, [_vp, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_float, C.c_int, _vp, C.c_size_t, _vp]),
    "maxstyle_tables": (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _f32p, _f32p, _f32p,
                                  _f32p, _f32p, C.c_int, _f32p, _f32p, _vp]),
    "maxstyle_p2p_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "maxstyle_tables_p2p": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      _vp, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _f32p, _f32p, _vp]),
    "maxstyle_rank_barrier": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "maxstyle_fwd_p2p": (C.c_int, [_vp, _vp, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _vp, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                   _vp, C.c_int, C.c_int, _vp, _vp, C.c_size_t, _vp]),
    "maxstyle_apply": (C.c_int, [_vp, _vp, _f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, _vp]),
    "maxstyle_fwd": (C.c_int, [_vp, _vp, _f32p, _f32p, _vp, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                               C.c_int, C.c_int, _vp, C.c_size_t, _vp]),
    "maxstyle_fwd_act": (C.c_int, [_vp, _vp, _f32p, _f32p, _vp, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_int, C.c_int, C.c_int, C.c_float, _vp, _vp, _vp, C.c_size_t, _vp]),
    "maxstyle_bwd_act": (C.c_int, [_vp, _vp, _vp, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _vp, _f32p, _f32p, _f32p,
                                   C.c_int, _f32p, _f32p, _f32p, C.POINTER(StepStruct),
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _vp, C.c_size_t, _vp]),
    "maxstyle_rescale": (C.c_int, [_vp, _vp, _vp, _vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "maxstyle_fwd_kernels": (C.c_int, [C.c_int] * 7),
    "maxstyle_fwd_geometry": (C.c_int, [C.c_int] * 6 + [C.POINTER(C.c_int)]),
    "maxstyle_workspace_status": (C.c_int, [_vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "maxstyle_bwd": (C.c_int, [_vp, _vp, _vp, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _vp, _f32p, _f32p, _f32p,
                               C.c_int, _f32p, _f32p, _f32p, C.POINTER(StepStruct),
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, _vp]),
    "maxstyle_ce2d_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "maxstyle_ce2d_fwd": (C.c_int, [_vp, _vp, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, _vp]),
    "maxstyle_ce2d_bwd": (C.c_int, [_vp, _vp, _f32p, _f32p, _f32p, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "maxstyle_ce2d_fwd_grad": (C.c_int, [_vp, _vp, _vp, C.c_int, _f32p, _f32p, _f32p, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, _vp, C.c_size_t, _vp]),
    "maxstyle_ce2d_scale": (C.c_int, [_vp, _f32p, C.c_int64, C.c_int, _vp]),
    "maxstyle_step": (C.c_int, [_f32p, _f32p, _f32p, C.POINTER(StepStruct), C.c_int, C.c_int, _vp]),
}

_lib = None
_lock = threading.Lock()


class MaxStyleLibraryError(RuntimeError):
    pass


def get_lib():
    """Load libmaxstyle_b200.so once.  Raises MaxStyleLibraryError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise MaxStyleLibraryError(
                    f"{LIB_PATH} is missing: the CUDA library is not built and maxstyle_b200 has no fallback path. "
                    "Run `python -m maxstyle_b200.build` (or __graft_entry__.build()).")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError here = header and library out of sync
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = get_lib().maxstyle_strerror(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")
