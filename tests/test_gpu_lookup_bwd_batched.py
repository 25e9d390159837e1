"""The batched lookup adjoint (`raftcorr_lookup_bwd_batched` + `raftcorr_build_bwd_folded`): every lookup of a backward
pass scattered in one launch, pooled levels folded on chip.  Autograd of `raft/core/corr.py:45-91` over the refinement
iterations of `raft/core/raft.py:141-144` and of the pooling cascade `corr.py:41-43`.

Checked against (a) the numpy oracle through the raw C ABI (level 0 of the gradient pyramid == fold(sum_i scatter_i)),
(b) the per-iteration `raftcorr_lookup_bwd` path of this library (red.global.add + fold passes; pinned on the reference's
autograd goldens in test_gpu_parity.py) through `CorrBlock` with `RAFTCORR_LOOKUP_BWD=immediate`, (c) the reference
recipe in fp64 autograd at the training shape.  Bit-level equality is not expected (summation order), so the bound is
a few fp32 ulps of the largest gradient for the exact mode and 1e-3 * max for the tensor-core modes.
"""
import ctypes
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _inputs(B, D, H, W, L, r, iters, seed, sigma=3.0, wild=False):
    g = torch.Generator(device=dev()).manual_seed(seed)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    ys, xs = torch.meshgrid(torch.arange(H, device=dev()), torch.arange(W, device=dev()), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1)
    coords, gouts = [], []
    flow = sigma * torch.randn((B, 2, H, W), generator=g, device=dev())
    for k in range(iters):
        flow = flow + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev())
        c = (grid + flow).contiguous()
        if wild and k % 2 == 0:  # windows hanging over / lying entirely outside every border, non-finite coordinates
            c[0, 0, 0, :4] = torch.tensor([-3.5, W + 2.25, 1e9, float("nan")], device=dev())[: min(4, W)]
            c[0, 1, 1, :4] = torch.tensor([-7.75, H - 0.5, -1e12, float("inf")], device=dev())[: min(4, W)]
            c[-1, :, -1, -1] = torch.tensor([W - 1.0, H - 1.0], device=dev())
        coords.append(c)
        gouts.append(torch.randn((B, L * (2 * r + 1) ** 2, H, W), generator=g, device=dev()))
    return f1, f2, coords, gouts


def _grads(f1, f2, coords, gouts, L, r, mode, path):
    import optical_flow_dnn_b200 as rc

    old = os.environ.get("RAFTCORR_LOOKUP_BWD")
    os.environ["RAFTCORR_LOOKUP_BWD"] = path
    try:
        t1 = f1.clone().requires_grad_()
        t2 = f2.clone().requires_grad_()
        blk = rc.CorrBlock(t1, t2, num_levels=L, radius=r, build_mode=mode)
        assert blk._defer_ok == (path == "batched")
        torch.autograd.backward([blk(c) for c in coords], gouts)
    finally:
        if old is None:
            del os.environ["RAFTCORR_LOOKUP_BWD"]
        else:
            os.environ["RAFTCORR_LOOKUP_BWD"] = old
    return t1.grad, t2.grad


def _rel(a, ref):
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max())


@pytest.mark.parametrize("shape", [
    # B, D, H, W, L, r, iters
    (2, 32, 23, 37, 4, 4, 5),     # ragged, odd sizes on every level
    (1, 16, 17, 50, 3, 3, 1),     # single lookup
    (2, 24, 16, 24, 1, 4, 3),     # no pooled levels
    (1, 32, 33, 65, 2, 1, 2),     # radius 1
    (2, 16, 20, 28, 4, 2, 19),    # more lookups than one launch takes: second launch accumulates
    (3, 8, 8, 9, 2, 4, 4),        # windows larger than the coarse level
])
def test_batched_equals_per_iteration_path_exact_mode(shape):
    B, D, H, W, L, r, iters = shape
    f1, f2, coords, gouts = _inputs(B, D, H, W, L, r, iters, seed=5, wild=True)
    a1, a2 = _grads(f1, f2, coords, gouts, L, r, "fp32", "immediate")
    b1, b2 = _grads(f1, f2, coords, gouts, L, r, "fp32", "batched")
    assert torch.isfinite(b1).all() and torch.isfinite(b2).all()
    e1, e2 = _rel(b1, a1), _rel(b2, a2)
    print(f"{shape}: dfmap1 {e1:.2e} dfmap2 {e2:.2e}")
    assert e1 <= 5e-6 and e2 <= 5e-6, (e1, e2)


@pytest.mark.parametrize("mode", ["tf32", "f16"])
def test_batched_equals_per_iteration_path_tensor_core_modes(mode):
    B, D, H, W, L, r, iters = 2, 64, 30, 44, 4, 4, 6
    f1, f2, coords, gouts = _inputs(B, D, H, W, L, r, iters, seed=9)
    a1, a2 = _grads(f1, f2, coords, gouts, L, r, mode, "immediate")
    b1, b2 = _grads(f1, f2, coords, gouts, L, r, mode, "batched")
    # both paths round the folded level 0 to TF32 before the same GEMMs; the sums differ by fp32 reordering only
    assert _rel(b1, a1) <= 2e-4 and _rel(b2, a2) <= 2e-4, (_rel(b1, a1), _rel(b2, a2))


def _ref_grads_fp64(f1, f2, coords, gouts, L, r):
    a = f1.double().requires_grad_()
    b = f2.double().requires_grad_()
    B, D, H, W = f1.shape
    vol = torch.matmul(a.view(B, D, H * W).transpose(1, 2), b.view(B, D, H * W)) / math.sqrt(D)
    levels = [vol.reshape(B * H * W, 1, H, W)]
    for _ in range(L - 1):
        levels.append(F.avg_pool2d(levels[-1], 2, stride=2))
    d = torch.linspace(-r, r, 2 * r + 1, device=dev(), dtype=torch.float64)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
    loss = 0
    for c, go in zip(coords, gouts):
        cc = c.double().permute(0, 2, 3, 1)
        out = []
        for l, lvl in enumerate(levels):
            hl, wl = lvl.shape[-2:]
            pts = cc.reshape(B * H * W, 1, 1, 2) / 2 ** l + delta
            grid = torch.stack([2 * pts[..., 0] / (wl - 1) - 1, 2 * pts[..., 1] / (hl - 1) - 1], dim=-1)
            out.append(F.grid_sample(lvl, grid, align_corners=True).view(B, H, W, -1))
        loss = loss + (torch.cat(out, dim=-1).permute(0, 3, 1, 2) * go.double()).sum()
    loss.backward()
    return a.grad, b.grad


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("tf32", 1e-3), ("f16", 1e-3)])
def test_batched_gradients_at_training_shape_vs_fp64_reference(mode, tol):
    """BASELINE.json config 3 shape (D = 256, 46x62, r = 4, L = 4), 12 refinement iterations, coords detached like
    raft.py:143 -- i.e. the pure batched path, as train.py runs it."""
    B, D, H, W, L, r, iters = 2, 256, 46, 62, 4, 4, 12
    f1, f2, coords, gouts = _inputs(B, D, H, W, L, r, iters, seed=33, sigma=4.0)
    e1, e2 = _ref_grads_fp64(f1, f2, coords, gouts, L, r)
    g1, g2 = _grads(f1, f2, coords, gouts, L, r, mode, "batched")
    r1, r2 = _rel(g1, e1), _rel(g2, e2)
    print(f"config-3 shape, {mode}: dfmap1 {r1:.2e} dfmap2 {r2:.2e} (rel to max|fp64 reference|)")
    assert r1 <= tol and r2 <= tol, (r1, r2)


def test_batched_level0_matches_oracle_through_the_c_abi():
    """Raw C ABI: level 0 written by raftcorr_lookup_bwd_batched == oracle fold(sum of oracle scatters), including the
    accumulate flag and the TF32 rounding flag."""
    from optical_flow_dnn_b200 import _lib
    from oracle import corr_oracle as O

    lib = _lib.lib()
    B, H, W, L, r, iters = 2, 13, 18, 3, 3, 4
    rng = np.random.default_rng(3)
    K = 2 * r + 1
    coords = [O.coords_grid(B, H, W) + rng.normal(0, 2.5, (B, 2, H, W)).astype(np.float32) for _ in range(iters)]
    gouts = [rng.normal(size=(B, L * K * K, H, W)).astype(np.float32) for _ in range(iters)]
    shapes = O.level_shapes(H, W, L)
    zero_pyr = [np.zeros((B, H * W, h, w), np.float64) for h, w in shapes]
    total = [np.zeros_like(z) for z in zero_pyr]
    for c, g in zip(coords, gouts):
        dp, _ = O.lookup_grad(zero_pyr, c.astype(np.float64), g.astype(np.float64), r)
        total = [t + d for t, d in zip(total, dp)]
    folded = total[-1]
    for l in range(L - 2, -1, -1):
        folded = total[l] + O.avg_pool2_adjoint(folded, *shapes[l])

    lay = _lib.make_layout(B, H, W, L, interleaved=2)   # the forward layout; gradients are row-major whatever it says
    glay = _lib.make_layout(B, H, W, L, interleaved=0)
    assert lib.raftcorr_lookup_bwd_batched_supported(ctypes.byref(lay), r) == 1
    tg = [torch.from_numpy(g).to(dev()) for g in gouts]
    tc = [torch.from_numpy(c).to(dev()) for c in coords]
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(sel, dpyr, accumulate, rnd):
        gp = (ctypes.c_void_p * len(sel))(*[tg[i].data_ptr() for i in sel])
        cp = (ctypes.c_void_p * len(sel))(*[tc[i].data_ptr() for i in sel])
        _lib.check(lib.raftcorr_lookup_bwd_batched(gp, cp, len(sel), ctypes.c_void_p(dpyr.data_ptr()), ctypes.byref(lay),
                                                   r, accumulate, rnd, 0, stream), "raftcorr_lookup_bwd_batched")

    def level0(dpyr):
        h, w, p = glay.h[0], glay.w[0], glay.pitch[0]
        return dpyr[: B * H * W * h * p].view(B, H * W, h, p)[..., :w].cpu().numpy()

    sentinel = 1234.5
    dpyr = torch.full((glay.total_floats,), sentinel, device=dev())
    run(range(iters), dpyr, 0, 0)
    got = level0(dpyr)
    scale = np.abs(folded).max()
    assert np.abs(got - folded).max() <= 2e-6 * scale
    if L > 1:  # the pooled levels of the buffer are not touched
        assert bool((dpyr[glay.offset[1]:] == sentinel).all())
    # two launches, the second accumulating, == one launch
    dpyr2 = torch.full((glay.total_floats,), sentinel, device=dev())
    run([0, 1], dpyr2, 0, 0)
    run([2, 3], dpyr2, 1, 0)
    assert np.abs(level0(dpyr2) - folded).max() <= 2e-6 * scale
    # TF32 rounding: 10 explicit mantissa bits, round to nearest
    dpyr3 = torch.empty(glay.total_floats, device=dev())
    run(range(iters), dpyr3, 0, 1)
    got3 = level0(dpyr3)
    assert (got3.view(np.uint32) & 0x1FFF == 0).all()
    assert np.abs(got3 - got).max() <= 2.0 ** -11 * 1.001 * scale
    # count outside 1..16 is an error, not a crash
    gp = (ctypes.c_void_p * 1)(tg[0].data_ptr())
    assert lib.raftcorr_lookup_bwd_batched(gp, gp, 17, ctypes.c_void_p(dpyr.data_ptr()), ctypes.byref(lay), r, 0, 0, 0,
                                           stream) != 0


def test_mixed_pass_lookups_plus_corr_pyramid_gradient():
    """A loss that uses lookups AND corr_pyramid: the queued lookups are flushed, the accumulator turns into the
    zero-filled scatter-add pyramid, and the result equals the per-iteration path."""
    import optical_flow_dnn_b200 as rc

    B, D, H, W, L, r, iters = 1, 32, 18, 26, 3, 3, 4
    f1, f2, coords, gouts = _inputs(B, D, H, W, L, r, iters, seed=12)
    res = {}
    for path in ("immediate", "batched"):
        os.environ["RAFTCORR_LOOKUP_BWD"] = path
        try:
            t1, t2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
            blk = rc.CorrBlock(t1, t2, num_levels=L, radius=r, build_mode="fp32")
            loss = sum((blk(c) * g).sum() for c, g in zip(coords[:2], gouts[:2]))
            loss = loss + sum((lv ** 2).sum() * 0.01 for lv in blk.corr_pyramid)
            loss = loss + sum((blk(c) * g).sum() for c, g in zip(coords[2:], gouts[2:]))
            loss.backward()
            res[path] = (t1.grad, t2.grad)
        finally:
            del os.environ["RAFTCORR_LOOKUP_BWD"]
    assert _rel(res["batched"][0], res["immediate"][0]) <= 5e-6
    assert _rel(res["batched"][1], res["immediate"][1]) <= 5e-6


def test_large_maps_fall_back_to_the_per_iteration_path():
    """A pixel's gradient pyramid that does not fit in shared memory: `supported` says no and CorrBlock keeps the
    per-iteration adjoint (still this library's kernels)."""
    from optical_flow_dnn_b200 import _lib

    lay = _lib.make_layout(1, 135, 240, 4, interleaved=2)
    assert _lib.lib().raftcorr_lookup_bwd_batched_supported(ctypes.byref(lay), 4) == 0
    lay = _lib.make_layout(8, 55, 128, 4, interleaved=2)
    assert _lib.lib().raftcorr_lookup_bwd_batched_supported(ctypes.byref(lay), 4) == 1


@pytest.mark.parametrize("cls_name", ["CorrBlock", "AlternateCorrBlock", "CorrBlock.from_encoder_head"])
def test_training_blocks_are_freed_by_refcount_not_by_the_cycle_collector(cls_name):
    """The autograd nodes must not reference the block (block -> pyramid -> grad_fn -> ctx -> block would keep every
    step's pyramid alive until Python's cycle collector runs: two cudaMalloc calls and 4.5 ms of host time per eager
    config-3 step before the core/facade split, scripts/eager_profile.py)."""
    import gc

    import optical_flow_dnn_b200 as rc

    B, D, H, W, L, r = 2, 64, 32, 48, 4, 4
    f1, f2, coords, gouts = _inputs(B, D, H, W, L, r, 3, seed=2)
    x = torch.randn((2 * B, 32, H, W), device=dev())
    w = (torch.randn((D, 32), device=dev()) / 6).requires_grad_()

    def step():
        a, b = f1.clone().requires_grad_(), f2.clone().requires_grad_()
        if cls_name == "CorrBlock":
            blk = rc.CorrBlock(a, b, L, r)
        elif cls_name == "AlternateCorrBlock":
            blk = rc.AlternateCorrBlock(a, b, L, r)
        else:
            blk = rc.CorrBlock.from_encoder_head(x, w, None, num_levels=L, radius=r)
        torch.autograd.backward([blk(c) for c in coords], gouts)
        w.grad = None

    step()
    gc.collect()
    torch.cuda.synchronize()
    gc.disable()
    try:
        base = torch.cuda.memory_allocated()
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        grown = torch.cuda.memory_allocated() - base
    finally:
        gc.enable()
    pyramid_bytes = B * H * W * H * W * 4
    assert grown < pyramid_bytes // 2, f"{grown} bytes still allocated after 4 steps without the cycle collector"


def test_second_backward_pass_over_a_retained_graph_and_an_aborted_pass():
    """The accumulator and the queue of lookup gradients are keyed to the autograd engine's graph task: a second
    backward over a retained graph starts fresh, and a pass that died half way leaks nothing into the next one."""
    import optical_flow_dnn_b200 as rc

    B, D, H, W, L, r, iters = 1, 32, 20, 28, 3, 3, 3
    f1, f2, coords, gouts = _inputs(B, D, H, W, L, r, iters, seed=8)
    t1, t2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
    class Boom(torch.autograd.Function):  # kills a backward pass
        @staticmethod
        def forward(ctx, x):
            return x.clone()

        @staticmethod
        def backward(ctx, g):
            raise RuntimeError("boom")

    blk = rc.CorrBlock(t1, t2, L, r, build_mode="fp32")
    out0 = blk(coords[0])
    bad = Boom.apply(out0)  # created BEFORE the other lookups: the engine runs it after their backward has queued gradients
    outs = [out0, blk(coords[1]), blk(coords[2])]
    torch.autograd.backward(outs, gouts, retain_graph=True)
    g1, g2 = t1.grad.clone(), t2.grad.clone()
    t1.grad = t2.grad = None
    with pytest.raises(RuntimeError, match="boom"):
        torch.autograd.backward([outs[2], outs[1], bad], [gouts[2], gouts[1], gouts[0]], retain_graph=True)
    assert len(blk._core._pending) > 0  # the aborted pass left queued gradients behind ...
    t1.grad = t2.grad = None
    torch.autograd.backward(outs, gouts)  # ... which a clean pass afterwards must not see
    assert _rel(t1.grad, g1) <= 1e-6 and _rel(t2.grad, g2) <= 1e-6
