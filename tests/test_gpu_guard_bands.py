"""Guard-band test: no kernel of the library writes outside the buffers it is given.

Every device buffer the host layer allocates while these operators run (outputs, workspaces, gradient
accumulators, operand copies, plans) is carved out of a larger allocation whose first and last 4 KiB hold a
sentinel byte pattern; after the work has finished every guard must still hold the pattern.  Ragged shapes (odd
maps, widths that are not a multiple of any tile, channel counts that need padding) are where an overrun would
hide: tiles clipped by hand, padded pitches, the interleaved rows' odd partner.  The results are compared with the
exact-fp32 build at the same time, so a kernel that stays inside its buffers by not doing its work fails too.

(compute-sanitizer is not available on the GPU pool; this is the bounds check the prompt's tooling note asks for
in its place.)
"""
import math

import pytest
import torch

from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu

GUARD = 4096        # bytes on either side; a multiple of every alignment the kernels ask for (TMA: 128)
PATTERN = 0xA5


class GuardedAllocator:
    """Replaces torch.empty / zeros / empty_like / zeros_like for CUDA tensors while active."""

    def __init__(self, monkeypatch):
        self.raw = []
        self._orig = {n: getattr(torch, n) for n in ("empty", "zeros", "empty_like", "zeros_like")}
        monkeypatch.setattr(torch, "empty", lambda *a, **k: self._new("empty", False, a, k))
        monkeypatch.setattr(torch, "zeros", lambda *a, **k: self._new("zeros", True, a, k))
        monkeypatch.setattr(torch, "empty_like", lambda t, **k: self._like("empty_like", False, t, k))
        monkeypatch.setattr(torch, "zeros_like", lambda t, **k: self._like("zeros_like", True, t, k))

    def _carve(self, shape, dtype, device, zero):
        item = self._orig["empty"]((), dtype=dtype).element_size()
        nbytes = math.prod(shape) * item
        raw = torch.full((nbytes + 2 * GUARD,), PATTERN, dtype=torch.uint8, device=device)
        self.raw.append(raw)
        body = raw[GUARD:GUARD + nbytes].view(dtype).view(shape)
        if zero:
            body.zero_()
        return body

    def _new(self, name, zero, args, kw):
        device = kw.get("device")
        if device is None or torch.device(device).type != "cuda" or kw.get("out") is not None:
            return self._orig[name](*args, **kw)
        shape = tuple(args[0]) if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)) else tuple(args)
        return self._carve(tuple(int(s) for s in shape), kw.get("dtype") or torch.float32, device, zero)

    def _like(self, name, zero, t, kw):
        if not t.is_cuda:
            return self._orig[name](t, **kw)
        return self._carve(tuple(t.shape), kw.get("dtype") or t.dtype, t.device, zero)

    def check(self):
        torch.cuda.synchronize()
        assert self.raw, "nothing was allocated through the guarded allocator"
        for i, raw in enumerate(self.raw):
            head, tail = raw[:GUARD], raw[-GUARD:]
            assert bool((head == PATTERN).all()), f"allocation {i} ({raw.numel() - 2 * GUARD} bytes): bytes BEFORE it were written"
            assert bool((tail == PATTERN).all()), f"allocation {i} ({raw.numel() - 2 * GUARD} bytes): bytes AFTER it were written"
        return len(self.raw)


def _inputs(B, D, H, W, seed):
    g = torch.Generator(device="cuda:0").manual_seed(seed)
    d = torch.device("cuda:0")
    f1 = torch.randn((B, D, H, W), generator=g, device=d)
    f2 = torch.randn((B, D, H, W), generator=g, device=d)
    coords = torch.from_numpy(O.coords_grid(B, H, W)).to(d) + 3.0 * torch.randn((B, 2, H, W), generator=g, device=d)
    # a few windows that hang over every border
    coords[:, :, 0, 0] = -2.5
    coords[:, 0, -1, -1] = W + 1.25
    coords[:, 1, -1, -1] = H + 0.75
    return f1, f2, coords


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# (B, D, H, W, L, r): odd maps, widths off every tile size, a channel count that needs K padding, five levels
SHAPES = [(2, 64, 23, 37, 4, 4), (1, 96, 46, 62, 4, 3), (3, 128, 17, 50, 3, 4), (1, 64, 33, 65, 5, 2)]


@pytest.mark.parametrize("shape", SHAPES)
def test_corr_block_stays_inside_its_buffers(shape, monkeypatch):
    import optical_flow_dnn_b200 as pkg

    B, D, H, W, L, r = shape
    f1, f2, coords = _inputs(B, D, H, W, 11)
    K2 = L * (2 * r + 1) ** 2
    w = torch.randn((B, K2, H, W), device=f1.device)
    with torch.no_grad():
        exact = pkg.CorrBlock(f1, f2, L, r, build_mode="fp32")(coords)
    alloc = GuardedAllocator(monkeypatch)
    for mode, tol in (("fp32", 0.0), ("tf32", 1e-3), ("f16", 1e-3), ("f16x3", 5e-5)):
        a1, a2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
        blk = pkg.CorrBlock(a1, a2, num_levels=L, radius=r, build_mode=mode)
        out = blk(coords)
        assert _rel(out.detach(), exact) <= tol, mode
        out2 = blk(coords + 0.37)          # a second iteration accumulates into the same gradient pyramid
        ((out * w).sum() + (out2 * w).sum()).backward()
        assert torch.isfinite(a1.grad).all() and torch.isfinite(a2.grad).all()
        pyr = blk.corr_pyramid
        assert len(pyr) == L and tuple(pyr[0].shape) == (B * H * W, 1, H, W)
    n = alloc.check()
    assert n >= 12


@pytest.mark.parametrize("shape", SHAPES[:3])
def test_alternate_corr_block_stays_inside_its_buffers(shape, monkeypatch):
    import optical_flow_dnn_b200 as pkg

    B, D, H, W, L, r = shape
    f1, f2, coords = _inputs(B, D, H, W, 12)
    w = torch.randn((B, L * (2 * r + 1) ** 2, H, W), device=f1.device)
    with torch.no_grad():
        exact = pkg.CorrBlock(f1, f2, L, r, build_mode="fp32")(coords)
    alloc = GuardedAllocator(monkeypatch)
    for mode, tol in (("fp32", 2e-5), ("tf32", 1e-3)):
        a1, a2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
        blk = pkg.AlternateCorrBlock(a1, a2, num_levels=L, radius=r, build_mode=mode)
        out = blk(coords)
        assert _rel(out.detach(), exact) <= tol, mode
        out2 = blk(coords - 0.6)
        ((out * w).sum() + (out2 * w).sum()).backward()
        assert torch.isfinite(a1.grad).all() and torch.isfinite(a2.grad).all()
    alloc.check()


def test_alt_cuda_corr_module_stays_inside_its_buffers(monkeypatch):
    from optical_flow_dnn_b200 import alt_cuda_corr

    B, D, H, W, r = 2, 64, 19, 29, 4
    f1, f2, coords = _inputs(B, D, H, W, 13)
    f1n, f2n = f1.permute(0, 2, 3, 1).contiguous(), f2.permute(0, 2, 3, 1).contiguous()
    ci = coords.permute(0, 2, 3, 1).reshape(B, 1, H, W, 2).contiguous()
    alloc = GuardedAllocator(monkeypatch)
    corr, = alt_cuda_corr.forward(f1n, f2n, ci, r)
    g = torch.randn(corr.shape, device=corr.device)
    d1, d2, dc = alt_cuda_corr.backward(f1n, f2n, ci, g, r)
    assert torch.isfinite(corr).all() and torch.isfinite(d1).all() and torch.isfinite(d2).all()
    assert float(dc.abs().max()) == 0.0
    alloc.check()


@pytest.mark.parametrize("shape", [(2, 64, 23, 37, 4, 4, 96), (1, 128, 46, 62, 4, 3, 80)])
def test_fused_ops_stay_inside_their_buffers(shape, monkeypatch):
    import optical_flow_dnn_b200 as pkg

    B, D, H, W, L, r, cout = shape
    d = torch.device("cuda:0")
    f1, f2, coords = _inputs(B, D, H, W, 14)
    g = torch.Generator(device=d).manual_seed(15)
    K2 = L * (2 * r + 1) ** 2
    cw = (torch.randn((cout, K2), generator=g, device=d) / 18.0)
    cb = torch.randn((cout,), generator=g, device=d)
    cin = 32
    hx = torch.randn((2 * B, cin, H, W), generator=g, device=d)
    hw = torch.randn((D, cin), generator=g, device=d) / cin ** 0.5
    hb = torch.randn((D,), generator=g, device=d) * 0.1
    flow = torch.randn((B, 2, H, W), generator=g, device=d)
    mask = torch.randn((B, 576, H, W), generator=g, device=d)
    with torch.no_grad():
        ref_c = torch.relu(torch.einsum("oc,bchw->bohw", cw, pkg.CorrBlock(f1, f2, L, r, build_mode="fp32")(coords))
                           + cb[None, :, None, None])
        fm = torch.einsum("oc,bchw->bohw", hw, hx) + hb[None, :, None, None]
        ref_h = pkg.CorrBlock(fm[:B].contiguous(), fm[B:].contiguous(), L, r, build_mode="fp32")(coords)

    alloc = GuardedAllocator(monkeypatch)
    # lookup -> convc1 -> ReLU, with its recomputing backward
    a1, a2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
    w_, b_ = cw.clone().requires_grad_(), cb.clone().requires_grad_()
    out = pkg.CorrBlock(a1, a2, num_levels=L, radius=r).lookup_convc1(coords, w_, b_)
    assert _rel(out.detach(), ref_c) <= 2e-3
    out.square().sum().backward()
    assert torch.isfinite(a1.grad).all() and torch.isfinite(w_.grad).all() and torch.isfinite(b_.grad).all()
    # encoder head -> build (both block types), with backward into x and the head's parameters
    for cls in (pkg.CorrBlock, pkg.AlternateCorrBlock):
        x_, hw_, hb_ = hx.clone().requires_grad_(), hw.clone().requires_grad_(), hb.clone().requires_grad_()
        o = cls.from_encoder_head(x_, hw_, hb_, num_levels=L, radius=r)(coords)
        assert _rel(o.detach(), ref_h) <= 3e-3, cls.__name__
        o.square().sum().backward()
        assert torch.isfinite(x_.grad).all() and torch.isfinite(hw_.grad).all() and torch.isfinite(hb_.grad).all()
    # convex 8x upsampling, forward and backward
    fl, mk = flow.clone().requires_grad_(), mask.clone().requires_grad_()
    up = pkg.upsample_flow(fl, mk)
    assert tuple(up.shape) == (B, 2, 8 * H, 8 * W)
    up.square().sum().backward()
    assert torch.isfinite(fl.grad).all() and torch.isfinite(mk.grad).all()
    alloc.check()


def test_guarded_allocator_detects_an_overrun(monkeypatch):
    """The checker itself: one float written just past a guarded tensor must be reported."""
    alloc = GuardedAllocator(monkeypatch)
    t = torch.empty((5, 7), dtype=torch.float32, device="cuda:0")
    z = torch.zeros(16, dtype=torch.uint8, device=torch.device("cuda:0"))
    assert float(z.sum()) == 0.0 and alloc.check() == 2
    flat = alloc.raw[0][GUARD:].view(torch.float32)      # body followed by its tail guard
    flat[t.numel()] = 1.0
    with pytest.raises(AssertionError, match="AFTER"):
        alloc.check()
