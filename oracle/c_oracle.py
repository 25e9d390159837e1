"""ctypes front-end of ``oracle/corr_oracle.c`` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_fp = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    so = os.path.join(_HERE, "libcorr_oracle.so")
    src = os.path.join(_HERE, "corr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libcorr_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(_fp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr_array(arrs):
    return (_fp * len(arrs))(*[_p(a) for a in arrs])


def num_threads():
    return int(lib().oracle_num_threads())


def corr_pyramid(fmap1, fmap2, num_levels=4):
    f1, f2 = _c(fmap1), _c(fmap2)
    B, D, H, W = f1.shape
    N = H * W
    v = np.empty((B, N, H, W), np.float32)
    lib().oracle_corr_volume(_p(f1), _p(f2), B, D, H, W, _p(v))
    pyr = [v]
    for _ in range(num_levels - 1):
        Hi, Wi = pyr[-1].shape[-2:]
        o = np.empty((B, N, Hi // 2, Wi // 2), np.float32)
        lib().oracle_avg_pool2(_p(pyr[-1]), ctypes.c_int64(B * N), Hi, Wi, _p(o))
        pyr.append(o)
    return pyr


def lookup(pyramid, coords, radius=4):
    pyr = [_c(p) for p in pyramid]
    co = _c(coords)
    B, _, H, W = co.shape
    L, K = len(pyr), 2 * radius + 1
    out = np.empty((B, L * K * K, H, W), np.float32)
    lib().oracle_lookup(_ptr_array(pyr), _p(co), B, H, W, L, radius, _p(out))
    return out


def lookup_grad(pyramid_shapes, coords, grad_out, radius=4):
    co, go = _c(coords), _c(grad_out)
    B, _, H, W = co.shape
    d = [np.zeros(s, np.float32) for s in pyramid_shapes]
    lib().oracle_lookup_grad(_p(co), _p(go), B, H, W, len(d), radius, _ptr_array(d))
    return d


def alt_lookup(fmap1, fmap2, coords, num_levels=4, radius=4):
    f1, f2, co = _c(fmap1), _c(fmap2), _c(coords)
    B, D, H, W = f1.shape
    levels = [f2]
    for _ in range(num_levels - 1):
        Hi, Wi = levels[-1].shape[-2:]
        o = np.empty((B, D, Hi // 2, Wi // 2), np.float32)
        lib().oracle_avg_pool2(_p(levels[-1]), ctypes.c_int64(B * D), Hi, Wi, _p(o))
        levels.append(o)
    K = 2 * radius + 1
    out = np.empty((B, num_levels * K * K, H, W), np.float32)
    lib().oracle_alt_lookup(_p(f1), _ptr_array(levels), _p(co), B, D, H, W, num_levels, radius, _p(out))
    return out
