"""Test-only stand-in for the *caller* of the correlation block: RAFT-small's refinement loop
(raft/core/raft.py:141-176) with its update operator (raft/core/update.py:183-201: SmallMotionEncoder
-> ConvGRU -> FlowHead), written from the architecture description, with deterministic seeded
weights so the same network can be rebuilt on the GPU box without shipping a checkpoint.

It exists to measure how a correlation-lookup error propagates through 12 recurrent iterations
(north_star: <= 0.01 px final-flow EPE delta).  tests/golden/make_golden_e2e.py proves, in the build
container, that this module is numerically identical to the reference's SmallUpdateBlock when given
the same weights, and records the reference's flows.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

HIDDEN, CONTEXT, RADIUS, LEVELS = 96, 64, 3, 4
COR_PLANES = LEVELS * (2 * RADIUS + 1) ** 2


class UpdateOperator(nn.Module):
    def __init__(self):
        super().__init__()
        # motion encoder (update.py:106-150)
        self.convc1 = nn.Conv2d(COR_PLANES, 96, 1)
        self.convf1 = nn.Conv2d(2, 64, 7, padding=3)
        self.convf2 = nn.Conv2d(64, 32, 3, padding=1)
        self.conv = nn.Conv2d(128, 80, 3, padding=1)
        # ConvGRU (update.py:57-79): input = context (64) + motion features (82)
        cin = HIDDEN + CONTEXT + 82
        self.convz = nn.Conv2d(cin, HIDDEN, 3, padding=1)
        self.convr = nn.Conv2d(cin, HIDDEN, 3, padding=1)
        self.convq = nn.Conv2d(cin, HIDDEN, 3, padding=1)
        # flow head (update.py:6-21)
        self.fh1 = nn.Conv2d(HIDDEN, 128, 3, padding=1)
        self.fh2 = nn.Conv2d(128, 2, 3, padding=1)

    def forward(self, net, inp, corr, flow):
        if isinstance(corr, tuple):  # (corr_fn, coords): the INTEGRATION.md 2c patch -- lookup fused with convc1 + ReLU
            corr_fn, coords = corr
            cor = corr_fn.lookup_convc1(coords, self.convc1.weight, self.convc1.bias)
        else:
            cor = F.relu(self.convc1(corr))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        mot = torch.cat([F.relu(self.conv(torch.cat([cor, flo], 1))), flow], 1)
        x = torch.cat([inp, mot], 1)
        hx = torch.cat([net, x], 1)
        z = torch.sigmoid(self.convz(hx))
        r = torch.sigmoid(self.convr(hx))
        q = torch.tanh(self.convq(torch.cat([r * net, x], 1)))
        net = (1 - z) * net + z * q
        return net, self.fh2(F.relu(self.fh1(net)))


# name in this module -> name in the reference's SmallUpdateBlock state_dict
REFERENCE_KEYS = {
    "convc1": "encoder.convc1", "convf1": "encoder.convf1", "convf2": "encoder.convf2", "conv": "encoder.conv",
    "convz": "gru.convz", "convr": "gru.convr", "convq": "gru.convq", "fh1": "flow_head.conv1", "fh2": "flow_head.conv2",
}


def seeded_operator(seed=2024):
    """Deterministic weights (CPU generator -> identical on every machine)."""
    op = UpdateOperator()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(op.named_parameters()):
            fan_in = p[0].numel() if p.dim() > 1 else p.numel()
            p.copy_(torch.randn(p.shape, generator=g) * (1.0 / fan_in ** 0.5 if p.dim() > 1 else 0.05))
    return op.eval()


def seeded_inputs(seed=7, B=1, D=128, H=24, W=32):
    """Feature maps with a real match structure: fmap2 is fmap1 displaced by a smooth flow + noise."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn((B, D, H + 8, W + 8), generator=g)
    base = F.avg_pool2d(base, 3, stride=1, padding=1) * 3.0
    f1 = base[:, :, 4:4 + H, 4:4 + W].contiguous()
    f2 = (base[:, :, 3:3 + H, 6:6 + W] + 0.1 * torch.randn((B, D, H, W), generator=g)).contiguous()
    net = torch.tanh(torch.randn((B, HIDDEN, H, W), generator=g))
    inp = torch.relu(torch.randn((B, CONTEXT, H, W), generator=g))
    return f1, f2, net, inp


def coords_grid(B, H, W, device):
    ys, xs = torch.meshgrid(torch.arange(H, device=device), torch.arange(W, device=device), indexing="ij")
    return torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1)


def refine(corr_cls, op, f1, f2, net, inp, iters=12, fused_convc1=False, **corr_kw):
    """raft.py:116-176 without the encoders / upsampling: returns the list of 1/8-resolution flows.
    ``fused_convc1``: hand the update operator (corr_fn, coords) instead of the looked-up features."""
    B, _, H, W = f1.shape
    corr_fn = corr_cls(f1, f2, num_levels=LEVELS, radius=RADIUS, **corr_kw)
    c0 = coords_grid(B, H, W, f1.device).to(f1.dtype)
    c1 = c0.clone()
    flows = []
    for _ in range(iters):
        c1 = c1.detach()
        corr = (corr_fn, c1) if fused_convc1 else corr_fn(c1).to(net.dtype)  # the reference ends in .float() (corr.py:91)
        net, delta = op(net, inp, corr, c1 - c0)
        c1 = c1 + delta
        flows.append(c1 - c0)
    return flows
