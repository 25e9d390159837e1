"""Host-side mirror of ``RAFT.upsample_flow`` (raft/core/raft.py:70-81): convex 8x upsampling of the 1/8-resolution
flow with the update block's mask -- SURVEY.md 8(f) rank 4, the per-iteration neighbour of the correlation lookup.

    upsample_flow(flow [B,C,H,W], mask [B,576,H,W]) -> [B,C,8H,8W]

Same arithmetic as the reference (softmax over the 9 neighbours, ``8 * flow``, zero padding at the border), one pass
over the mask in libraftcorr_b200 instead of softmax / unfold / multiply / sum / permute; differentiable w.r.t. both
inputs.  ``C`` may be up to 4: the uncertainty model upsamples flow and flow variance with the same mask
(raft.py:165-168) -- ``torch.cat`` them and split the result to read the mask once.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check
from .corr import _ptr, _stream


def _check(flow, mask):
    for name, t in (("flow", flow), ("mask", mask)):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4):
            raise RuntimeError(f"upsample_flow: {name} must be a 4-D float32 CUDA tensor (this package has no CPU path)")
    B, C, H, W = flow.shape
    if tuple(mask.shape) != (B, 576, H, W) or mask.device != flow.device:
        raise RuntimeError(f"upsample_flow: mask must be [B, 576, H, W] = {(B, 576, H, W)} on {flow.device}, got {tuple(mask.shape)}")
    if not 1 <= C <= 4:
        raise RuntimeError(f"upsample_flow: 1..4 channels (got {C})")


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, mask):
        ctx.save_for_backward(flow, mask)
        return _forward(flow, mask)

    @staticmethod
    def backward(ctx, gout):
        flow, mask = ctx.saved_tensors
        B, C, H, W = flow.shape
        dev = flow.device
        dmask = torch.empty_like(mask)
        dflow = torch.empty_like(flow)
        gout = gout.contiguous()  # bound to a local: must outlive the launch
        check(_lib.lib().raftcorr_upsample_flow_bwd(_ptr(flow), _ptr(mask), _ptr(gout), B, C, H, W, _ptr(dmask),
                                                    _ptr(dflow), dev.index, _stream(dev)), "raftcorr_upsample_flow_bwd")
        return dflow, dmask


def _forward(flow, mask):
    B, C, H, W = flow.shape
    dev = flow.device
    out = torch.empty((B, C, 8 * H, 8 * W), dtype=torch.float32, device=dev)
    check(_lib.lib().raftcorr_upsample_flow_fwd(_ptr(flow), _ptr(mask), B, C, H, W, _ptr(out), dev.index, _stream(dev)),
          "raftcorr_upsample_flow_fwd")
    return out


def upsample_flow(flow, mask):
    """Drop-in for ``RAFT.upsample_flow(self, flow, mask)`` (raft/core/raft.py:70-81)."""
    _check(flow, mask)
    flow, mask = flow.contiguous(), mask.contiguous()
    if torch.is_grad_enabled() and (flow.requires_grad or mask.requires_grad):
        return _UpsampleFn.apply(flow, mask)
    return _forward(flow, mask)
