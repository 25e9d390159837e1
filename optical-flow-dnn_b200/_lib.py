"""ctypes binding of libraftcorr_b200.so (the C ABI declared in include/raftcorr.h).

There is deliberately NO fallback: if the CUDA library is missing this module raises, and every
operator in this package fails loudly rather than silently running somewhere else.
"""
from __future__ import annotations

import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libraftcorr_b200.so")
MAX_LEVELS = 8
BUILD_FP32_SIMT = 0
BUILD_TF32_TC = 1
BUILD_F16_TC = 2
BUILD_F16X3_TC = 3
LOOKUP_PLAN_BYTES = 4096
LOOKUP_BWD_MAX_BATCH = 16


class RaftCorrError(RuntimeError):
    """Raised when a libraftcorr_b200 entry point returns a non-zero status."""


class PyramidLayout(ctypes.Structure):
    """Mirror of ``raftcorr_pyramid_layout`` (include/raftcorr.h)."""

    _fields_ = [
        ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("L", ctypes.c_int32),
        ("h", ctypes.c_int32 * MAX_LEVELS), ("w", ctypes.c_int32 * MAX_LEVELS),
        ("pitch", ctypes.c_int32 * MAX_LEVELS), ("offset", ctypes.c_int64 * MAX_LEVELS),
        ("total_floats", ctypes.c_int64), ("interleaved", ctypes.c_int32), ("reserved_", ctypes.c_int32),
    ]


_c_int, _c_void_p, _c_size_t, _c_i64 = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64
_LAYP = ctypes.POINTER(PyramidLayout)

# name -> (restype, argtypes); kept in lock-step with include/raftcorr.h (tests/test_abi.py checks it)
SIGNATURES = {
    "raftcorr_abi_version": (_c_int, []),
    "raftcorr_last_error": (ctypes.c_char_p, []),
    "raftcorr_set_l2_fetch_granularity": (_c_int, [_c_int, _c_int, ctypes.POINTER(_c_int)]),
    "raftcorr_pyramid_layout_init": (_c_int, [_LAYP, _c_int, _c_int, _c_int, _c_int, _c_int]),
    "raftcorr_build_fwd_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "raftcorr_build_fwd": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _LAYP, _c_int, _c_void_p, _c_size_t,
                                     _c_int, _c_void_p]),
    "raftcorr_fnet_head_supported": (_c_int, [_c_int] * 4),
    "raftcorr_fnet_head_build_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "raftcorr_fnet_head_build_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _LAYP, _c_void_p,
                                               _c_size_t, _c_int, _c_void_p]),
    "raftcorr_fnet_head_alt_workspace_bytes": (_c_size_t, [_c_int] * 2),
    "raftcorr_fnet_head_alt_pyramid": (_c_int, [_c_void_p] * 3 + [_c_int] * 6 + [_c_void_p, _c_void_p, _c_void_p, _c_size_t,
                                                                                   _c_int, _c_void_p]),
    "raftcorr_lookup_fwd": (_c_int, [_c_void_p, _LAYP, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p]),
    "raftcorr_lookup_plan_init": (_c_int, [_c_void_p, _c_void_p, _LAYP, _c_int, _c_int]),
    "raftcorr_lookup_fwd_planned": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p]),
    "raftcorr_convc1_weight_bytes": (_c_size_t, [_c_int] * 3),
    "raftcorr_convc1_pack_weight": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_void_p]),
    "raftcorr_lookup_convc1_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int,
                                             _c_void_p]),
    "raftcorr_upsample_flow_fwd": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_void_p]),
    "raftcorr_upsample_flow_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                             _c_void_p, _c_int, _c_void_p]),
    "raftcorr_lookup_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _LAYP, _c_int, _c_void_p, _c_int,
                                      _c_void_p]),
    "raftcorr_build_bwd_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "raftcorr_build_bwd": (_c_int, [_c_void_p, _LAYP, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int,
                                     _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "raftcorr_lookup_bwd_batched_supported": (_c_int, [_LAYP, _c_int]),
    "raftcorr_lookup_bwd_batched": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _LAYP, _c_int, _c_int, _c_int, _c_int,
                                              _c_void_p]),
    "raftcorr_build_bwd_wants_tf32": (_c_int, [_c_int, _c_int]),
    "raftcorr_build_bwd_folded": (_c_int, [_c_void_p, _LAYP, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int,
                                            _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "raftcorr_alt_forward": (_c_int, [_c_void_p] * 3 + [_c_int] * 8 + [_c_void_p, _c_int, _c_void_p]),
    "raftcorr_alt_backward": (_c_int, [_c_void_p] * 4 + [_c_int] * 8 + [_c_void_p] * 3 + [_c_int, _c_void_p]),
    "raftcorr_alt_level_offset": (_c_i64, [_c_int] * 5),
    "raftcorr_alt_pyramid": (_c_int, [_c_void_p, _c_void_p] + [_c_int] * 6 + [_c_void_p, _c_void_p, _c_int, _c_void_p]),
    "raftcorr_alt_lookup_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "raftcorr_alt_lookup_fwd": (_c_int, [_c_void_p] * 3 + [_c_int] * 7 + [_c_void_p, _c_void_p, _c_size_t, _c_int,
                                                                            _c_void_p]),
    "raftcorr_alt_f16_bytes": (_c_size_t, [_c_int] * 5),
    "raftcorr_alt_f16_prepare": (_c_int, [_c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p, _c_int, _c_void_p]),
    "raftcorr_alt_lookup_fwd_f16": (_c_int, [_c_void_p, _c_void_p] + [_c_int] * 6 + [_c_void_p, _c_void_p, _c_size_t, _c_int,
                                                                                      _c_void_p]),
    "raftcorr_alt_lookup_bwd": (_c_int, [_c_void_p] * 4 + [_c_int] * 6 + [_c_void_p, _c_void_p, _c_int, _c_void_p]),
    "raftcorr_alt_grad_finish": (_c_int, [_c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p, _c_void_p, _c_int,
                                                                                  _c_void_p]),
}

_LIB = None


def lib():
    """Load the shared library once.  Raises if it has not been built (no CPU / eager fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RaftCorrError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This package has no fallback path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI drifted
            fn.restype, fn.argtypes = res, args
        if handle.raftcorr_abi_version() != 1:
            raise RaftCorrError("libraftcorr_b200.so ABI version mismatch")
        _LIB = handle
    return _LIB


def check(status: int, what: str):
    if status != 0:
        msg = lib().raftcorr_last_error()
        raise RaftCorrError(f"{what}: {msg.decode() if msg else 'unknown error'}")


def make_layout(B: int, H: int, W: int, L: int, interleaved: int = 0) -> PyramidLayout:
    """``interleaved``: 0 plain rows, 1 row pairs, 2 row quads (include/raftcorr.h)."""
    lay = PyramidLayout()
    check(lib().raftcorr_pyramid_layout_init(ctypes.byref(lay), B, H, W, L, int(interleaved)), "pyramid layout")
    return lay
