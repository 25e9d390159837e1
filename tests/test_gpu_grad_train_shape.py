"""Gradient parity at the sizes the reference trains at (`raft/train.py:71,118-128`; BASELINE.json config 3: D = 256,
368x496 -> 46x62 feature maps, r = 4, L = 4; RAFT-small: D = 128, r = 3), several refinement iterations accumulated.

Reference: the reference's own recipe (`corr.py:93-116` matmul / sqrt(D), `:41-43` avg_pool2d cascade, `:45-91` +
`utils.py:57-85` grid_sample lookup) evaluated in FP64 on the device with torch autograd -- the "true" gradients -- plus
(a) the exact-fp32 build mode of this library (pinned on the reference goldens) and (b) an independent fp64 contraction
of sampled rows/columns of the level-0 volume gradient (dF1[m] = sum_n dV0[m,n] F2[n] / sqrt(D), dF2[n] = sum_m ...).
In the backward GEMMs D is the MMA N dimension: D = 32 (goldens), 128 and 256 are different instruction shapes.

Tolerance: max|delta| <= 1e-3 * max|ref| per tensor; d/dcoords <= 2 x the forward tolerance of the mode.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = {"fp32": 2e-5, "tf32": 1e-3, "f16": 1e-3}
GRAD_TOL = 1e-3


def dev():
    return torch.device("cuda:0")


def _ref_pyramid(f1, f2, L):
    B, D, H, W = f1.shape
    vol = torch.matmul(f1.view(B, D, H * W).transpose(1, 2), f2.view(B, D, H * W)) / math.sqrt(D)
    levels = [vol.reshape(B * H * W, 1, H, W)]
    for _ in range(L - 1):
        levels.append(F.avg_pool2d(levels[-1], 2, stride=2))
    return levels


def _ref_lookup(levels, coords, r):
    B, _, H, W = coords.shape
    c = coords.permute(0, 2, 3, 1)
    d = torch.linspace(-r, r, 2 * r + 1, device=coords.device, dtype=coords.dtype)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
    out = []
    for l, lvl in enumerate(levels):
        hl, wl = lvl.shape[-2:]
        pts = c.reshape(B * H * W, 1, 1, 2) / 2 ** l + delta
        gx = 2 * pts[..., 0] / (wl - 1) - 1
        gy = 2 * pts[..., 1] / (hl - 1) - 1
        s = F.grid_sample(lvl, torch.stack([gx, gy], dim=-1), align_corners=True)
        out.append(s.view(B, H, W, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous()


def _case(B, D, H, W, L, r, iters, seed):
    g = torch.Generator(device=dev()).manual_seed(seed)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    ys, xs = torch.meshgrid(torch.arange(H, device=dev()), torch.arange(W, device=dev()), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1)
    coords, gouts = [], []
    flow = F.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev()), 5, stride=1, padding=2)
    for _ in range(iters):
        flow = flow + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev())
        coords.append((grid + flow).contiguous())
        gouts.append(torch.randn((B, L * (2 * r + 1) ** 2, H, W), generator=g, device=dev()))
    return f1, f2, coords, gouts


def _reference_grads(f1, f2, coords, gouts, L, r):
    a = f1.double().requires_grad_()
    b = f2.double().requires_grad_()
    levels = _ref_pyramid(a, b, L)
    levels[0].retain_grad()
    c1 = coords[1].double().requires_grad_()
    loss = 0
    for k, (c, go) in enumerate(zip(coords, gouts)):
        loss = loss + (_ref_lookup(levels, c1 if k == 1 else c.double(), r) * go.double()).sum()
    loss.backward()
    return a.grad, b.grad, c1.grad, levels[0].grad  # levels[0].grad = TOTAL gradient w.r.t. the level-0 volume


def _rel(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("shape", ["config3_D256_r4", "small_D128_r3"])
@pytest.mark.parametrize("mode", ["fp32", "tf32", "f16"])
def test_gradients_at_training_shape(shape, mode):
    import optical_flow_dnn_b200 as rc

    B, H, W, L, iters = 2, 46, 62, 4, 3
    D, r = (256, 4) if shape == "config3_D256_r4" else (128, 3)
    f1, f2, coords, gouts = _case(B, D, H, W, L, r, iters, seed=21)
    e1, e2, ec, dv0 = _reference_grads(f1, f2, coords, gouts, L, r)

    # independent fp64 contraction of sampled rows / columns of the level-0 volume gradient (checks the reference chain)
    N = H * W
    dv0 = dv0.view(B, N, N)
    idx = torch.tensor([0, 1, 61, 62, 1000, 1425, N - 63, N - 1], device=dev())
    a, b = f1.double().view(B, D, N), f2.double().view(B, D, N)
    s1 = torch.einsum("bmn,bdn->bdm", dv0[:, idx, :], b) / math.sqrt(D)
    s2 = torch.einsum("bmn,bdm->bdn", dv0[:, :, idx], a) / math.sqrt(D)
    assert float((s1 - e1.view(B, D, N)[:, :, idx]).abs().max()) <= 1e-9 * float(e1.abs().max())
    assert float((s2 - e2.view(B, D, N)[:, :, idx]).abs().max()) <= 1e-9 * float(e2.abs().max())

    t1 = f1.clone().requires_grad_()
    t2 = f2.clone().requires_grad_()
    tc = coords[1].clone().requires_grad_()
    blk = rc.CorrBlock(t1, t2, num_levels=L, radius=r, build_mode=mode)
    loss = 0
    for k, (c, go) in enumerate(zip(coords, gouts)):
        loss = loss + (blk(tc if k == 1 else c) * go).sum()
    loss.backward()
    r1, r2, rc_ = _rel(t1.grad, e1), _rel(t2.grad, e2), _rel(tc.grad, ec)
    print(f"{shape} {mode}: dfmap1 {r1:.2e}  dfmap2 {r2:.2e}  dcoords {rc_:.2e}  (rel to max|fp64 reference|)")
    assert r1 <= GRAD_TOL, ("dfmap1", r1)
    assert r2 <= GRAD_TOL, ("dfmap2", r2)
    assert rc_ <= 2 * TOL[mode], ("dcoords", rc_)
    # sampled rows/columns of OUR gradients against the fp64 contraction of the reference's dV0
    assert float((t1.grad.view(B, D, N)[:, :, idx].double() - s1).abs().max()) <= GRAD_TOL * float(e1.abs().max())
    assert float((t2.grad.view(B, D, N)[:, :, idx].double() - s2).abs().max()) <= GRAD_TOL * float(e2.abs().max())


@pytest.mark.parametrize("mode", ["tf32", "f16"])
def test_tensor_core_gradients_match_exact_fp32_mode(mode):
    """Same graph in the exact-fp32 build mode (itself pinned on the reference goldens): B = 3 so the batch loop of the
    backward GEMMs is exercised past a pair."""
    import optical_flow_dnn_b200 as rc

    B, D, H, W, L, r, iters = 3, 256, 46, 62, 4, 4, 3
    f1, f2, coords, gouts = _case(B, D, H, W, L, r, iters, seed=33)
    grads = {}
    for m in ("fp32", mode):
        t1 = f1.clone().requires_grad_()
        t2 = f2.clone().requires_grad_()
        blk = rc.CorrBlock(t1, t2, num_levels=L, radius=r, build_mode=m)
        loss = 0
        for c, go in zip(coords, gouts):
            loss = loss + (blk(c) * go).sum()
        loss.backward()
        grads[m] = (t1.grad, t2.grad)
    for name, got, ref in (("dfmap1", grads[mode][0], grads["fp32"][0]), ("dfmap2", grads[mode][1], grads["fp32"][1])):
        err = _rel(got, ref.double())
        print(f"{mode} vs fp32 mode, {name}: {err:.2e}")
        assert err <= GRAD_TOL, (name, err)


@pytest.mark.parametrize("mode", ["fp32", "tf32", "f16"])
def test_corr_pyramid_and_static_corr_are_differentiable(mode):
    """The reference's `CorrBlock.corr` (corr.py:93-116) and `corr_pyramid` list are ordinary differentiable tensors.
    Here the pyramid buffer is row-pair interleaved in the tensor-core modes while gradient pyramids are row-major: a
    loss on `corr_pyramid` levels, on `CorrBlock.corr`, and on a MIX of lookups and levels must all give the reference
    gradients."""
    import optical_flow_dnn_b200 as rc

    B, D, H, W, L, r = 2, 64, 19, 22, 3, 3
    f1, f2, coords, gouts = _case(B, D, H, W, L, r, 2, seed=5)
    g = torch.Generator(device=dev()).manual_seed(9)

    def ref_levels(a, b):
        return _ref_pyramid(a, b, L)

    # (1) static corr
    a, b = f1.double().requires_grad_(), f2.double().requires_grad_()
    v = ref_levels(a, b)[0].view(B, H, W, 1, H, W)
    wv = torch.randn(v.shape, generator=g, device=dev())
    (v * wv.double()).sum().backward()
    t1, t2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
    ours = rc.CorrBlock.corr(t1, t2, build_mode=mode)
    assert ours.shape == v.shape and _rel(ours, v.detach()) <= TOL[mode]
    (ours * wv).sum().backward()
    assert _rel(t1.grad, a.grad) <= GRAD_TOL and _rel(t2.grad, b.grad) <= GRAD_TOL

    # (2) a loss on pyramid levels only, (3) lookups + levels mixed, (4) the same mixed loss again after an ABORTED pass
    for mixed in (False, True, True):
        a, b = f1.double().requires_grad_(), f2.double().requires_grad_()
        lv = ref_levels(a, b)
        ws = [torch.randn(x.shape, generator=g, device=dev()) for x in lv]
        loss = sum((x * w.double()).sum() for x, w in zip(lv, ws))
        if mixed:
            loss = loss + sum((_ref_lookup(lv, c.double(), r) * go.double()).sum() for c, go in zip(coords, gouts))
        loss.backward()
        t1, t2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
        blk = rc.CorrBlock(t1, t2, num_levels=L, radius=r, build_mode=mode)
        levels = blk.corr_pyramid
        loss = 0
        if mixed:
            for c, go in zip(coords, gouts):
                loss = loss + (blk(c) * go).sum()
        loss = loss + sum((x * w).sum() for x, w in zip(levels, ws))
        if mixed:
            # an aborted backward (exception after the first lookup backward has opened the accumulator) must not leak
            class _Boom(torch.autograd.Function):
                @staticmethod
                def forward(ctx, x):
                    return x.clone()

                @staticmethod
                def backward(ctx, gx):
                    raise ValueError("boom")

            bad = (blk(coords[0]) * gouts[0]).sum() + _Boom.apply(t1).sum() * 0
            with pytest.raises(ValueError):
                bad.backward()
            t1.grad = t2.grad = None
        loss.backward()
        e1, e2 = _rel(t1.grad, a.grad), _rel(t2.grad, b.grad)
        print(f"{mode} mixed={mixed}: dfmap1 {e1:.2e} dfmap2 {e2:.2e}")
        assert e1 <= GRAD_TOL and e2 <= GRAD_TOL, (mixed, e1, e2)
