"""Mirror of the reference's native extension module ``alt_cuda_corr``
(raft/alt_cuda_corr/correlation.cpp:51-54): same function names, argument order, tensor layouts,
return convention (a list of tensors) and error behaviour (RuntimeError for non-CUDA or
non-contiguous inputs, correlation.cpp:19-21) -- backed by libraftcorr_b200.

    forward(fmap1[B,H1,W1,C], fmap2[B,H2,W2,C], coords[B,N,H1,W1,2], radius) -> [corr[B,N,(2r+1)^2,H1,W1]]
    backward(fmap1, fmap2, coords, corr_grad[B,N,(2r+1)^2,H1,W1], radius)
        -> [fmap1_grad[B,H1,W1,C], fmap2_grad[B,H2,W2,C], coords_grad[B,N,H1,W1,2] (zeros)]

A maintainer of the reference makes ``import alt_cuda_corr`` (raft/core/corr.py:7) resolve to this
module; see INTEGRATION.md.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check


def _check_input(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32 (the reference kernel is hard-instantiated for float, "
                           "correlation_kernel.cu:278)")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _dims(fmap1, fmap2, coords):
    B, H1, W1, C = fmap1.shape
    B2, H2, W2, C2 = fmap2.shape
    Bc, N, Hc, Wc, two = coords.shape
    if (B2, C2) != (B, C) or (Bc, Hc, Wc, two) != (B, H1, W1, 2):
        raise RuntimeError("alt_cuda_corr: inconsistent tensor shapes")
    return B, N, H1, W1, H2, W2, C


def forward(fmap1, fmap2, coords, radius):
    for t, n in ((fmap1, "fmap1"), (fmap2, "fmap2"), (coords, "coords")):
        _check_input(t, n)
    B, N, H1, W1, H2, W2, C = _dims(fmap1, fmap2, coords)
    rd = 2 * radius + 1
    dev = fmap1.device
    corr = torch.empty((B, N, rd * rd, H1, W1), dtype=torch.float32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(_lib.lib().raftcorr_alt_forward(_p(fmap1), _p(fmap2), _p(coords), B, N, H1, W1, H2, W2, C, radius, _p(corr),
                                          dev.index, stream), "raftcorr_alt_forward")
    return [corr]


def backward(fmap1, fmap2, coords, corr_grad, radius):
    for t, n in ((fmap1, "fmap1"), (fmap2, "fmap2"), (coords, "coords"), (corr_grad, "corr_grad")):
        _check_input(t, n)
    B, N, H1, W1, H2, W2, C = _dims(fmap1, fmap2, coords)
    dev = fmap1.device
    fmap1_grad = torch.zeros_like(fmap1)
    fmap2_grad = torch.zeros_like(fmap2)
    coords_grad = torch.zeros_like(coords)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(_lib.lib().raftcorr_alt_backward(_p(fmap1), _p(fmap2), _p(coords), _p(corr_grad), B, N, H1, W1, H2, W2, C,
                                           radius, _p(fmap1_grad), _p(fmap2_grad), _p(coords_grad), dev.index, stream),
          "raftcorr_alt_backward")
    return [fmap1_grad, fmap2_grad, coords_grad]
