"""CPU port of the reference CorrBlock on the same PyTorch CPU operators the reference calls --
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

Used by ``bench.py`` as the timed CPU baseline (``cpu_baseline.kind = "port"``, and the
``--impl reference`` arm): the reference's hot path IS three ATen calls -- ``torch.matmul``
(raft/core/corr.py:112), ``F.avg_pool2d`` (:42) and ``F.grid_sample`` (raft/core/utils/utils.py:78) --
so the fairest CPU baseline runs exactly those operators, multi-threaded (MKL/OpenMP), in fp32.
The reference's own Python cannot travel to the GPU box, hence a port; it is pinned against the
golden vectors in tests/test_oracle_golden.py like the other oracles.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class CpuCorrBlock:
    """Same constructor / call surface as the reference class (raft/core/corr.py:18,45)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        B, D, H, W = fmap1.shape
        self.num_levels, self.radius = num_levels, radius
        vol = torch.bmm(fmap1.reshape(B, D, H * W).transpose(1, 2), fmap2.reshape(B, D, H * W))
        vol.mul_(1.0 / math.sqrt(D))  # corr.py:116
        level = vol.reshape(B * H * W, 1, H, W)  # corr.py:35
        self.corr_pyramid = [level]
        for _ in range(num_levels - 1):
            level = F.avg_pool2d(level, 2, stride=2)  # corr.py:42
            self.corr_pyramid.append(level)
        r = radius
        off = torch.arange(-r, r + 1, dtype=fmap1.dtype)
        # window[i, j] = (offset_i, offset_j): first index added to x, second to y (corr.py:66-68,78)
        self._window = torch.stack(torch.meshgrid(off, off, indexing="ij"), dim=-1).reshape(1, 2 * r + 1, 2 * r + 1, 2)

    def __call__(self, coords):
        B, _, H, W = coords.shape
        centre = coords.permute(0, 2, 3, 1).reshape(B * H * W, 1, 1, 2)
        feats = []
        for l, level in enumerate(self.corr_pyramid):
            Hl, Wl = level.shape[-2:]
            pts = centre / (2 ** l) + self._window
            gx = 2 * pts[..., 0] / (Wl - 1) - 1  # utils.py:74-75
            gy = 2 * pts[..., 1] / (Hl - 1) - 1
            samp = F.grid_sample(level, torch.stack([gx, gy], dim=-1), mode="bilinear", align_corners=True)
            feats.append(samp.reshape(B, H, W, -1))
        return torch.cat(feats, dim=-1).permute(0, 3, 1, 2).contiguous().float()
