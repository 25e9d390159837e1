"""SURVEY.md 8(f) rank 2 on the GPU: the feature encoder's output head (1x1 ``conv2`` + split + ``.float()``) fused in
front of the tensor-core build (`raftcorr_fnet_head_build_fwd`, called through the C ABI) against

  (1) golden vectors from the reference encoders' own ``conv2`` (input captured by a forward hook inside
      ``self.fnet([image_1, image_2])``) + the reference CorrBlock on its output (tests/golden/make_golden_fnet_head.py;
      raft/core/extractor.py:231,239 / :341,349, raft/core/raft.py:110-119),
  (2) the CPU oracle on seeded shapes (ragged pixel tiles, partial channel blocks, padded D, no bias), operand
      buffers checked element by element,
  (3) at BASELINE.json's full size: ``CorrBlock(*split(conv2(x)))`` with the convolution in fp64 on the device,
  (4) autograd: gradients w.r.t. x, weight and bias against torch autograd through the unfused fp32 path.

Tolerances: operand buffers max|delta| <= 1e-3 * max|ref| (one TF32 product, fp32 accumulate, result rounded to TF32:
measured ~3e-4).  Volume and lookups 1.5e-3 * max|ref|: two TF32 stages stack here (the head, then north_star's 1e-3
budget for the volume built from TF32-rounded maps); measured 3e-4 .. 6e-4.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu
TOL_OPERAND = 1e-3
TOL = 1.5e-3


def dev():
    return torch.device("cuda:0")


def rel_err(a, ref):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    ref = ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else ref
    return float(np.abs(a.astype(np.float64) - ref.astype(np.float64)).max() / max(np.abs(ref).max(), 1e-30))


@pytest.fixture(scope="module")
def pkg():
    import optical_flow_dnn_b200 as m

    return m


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def test_golden_reference_encoder_head(pkg, golden_fnet_head):
    g = golden_fnet_head
    B, Cin, D, H, W, L, r = [int(v) for v in g["meta"]]
    with torch.no_grad():
        blk = pkg.CorrBlock.from_encoder_head(t(g["x"]), t(g["weight"]), t(g["bias"]), num_levels=L, radius=r)
        assert blk.num_levels == L and blk.radius == r
        v0 = blk.corr_pyramid[0].reshape(B, H * W, H, W)
        assert rel_err(v0, g["pyr0"]) <= TOL
        for k in range(3):
            out = blk(t(g[f"coords{k}"]))
            assert tuple(out.shape) == (B, L * (2 * r + 1) ** 2, H, W) and out.is_contiguous()
            assert rel_err(out, g[f"out64_{k}"]) <= TOL, f"coords set {k}"


def _operands(pkg, x, w, b, B, Cin, D, H, W, L=4):
    """Call the C ABI directly and return the two operand buffers the head wrote (+ the pyramid)."""
    from optical_flow_dnn_b200 import _lib

    lib = _lib.lib()
    lay = _lib.make_layout(B, H, W, L, interleaved=2)
    Dp = (D + 31) // 32 * 32
    ws_bytes = int(lib.raftcorr_fnet_head_build_workspace_bytes(B, Cin, D, H, W))
    ops_bytes = int(lib.raftcorr_build_fwd_workspace_bytes(B, D, H, W, _lib.BUILD_TF32_TC))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev())
    pyr = torch.empty(lay.total_floats, dtype=torch.float32, device=dev())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda z: ctypes.c_void_p(z.data_ptr()) if z is not None else ctypes.c_void_p(0)
    _lib.check(lib.raftcorr_fnet_head_build_fwd(p(x), p(w), p(b), Cin, D, p(pyr), ctypes.byref(lay), p(ws), ws_bytes, 0, st),
               "raftcorr_fnet_head_build_fwd")
    torch.cuda.synchronize()
    f = ws[:ops_bytes].view(torch.float32)
    half = ops_bytes // 8
    f1n = f[: B * H * W * Dp].view(B, H * W, Dp)
    f2n = f[half: half + B * H * W * Dp].view(B, H * W, Dp)
    return f1n, f2n, pyr


SEEDED = [
    # B, Cin, D, H, W, bias
    (1, 128, 256, 16, 20, True),    # RAFT-basic head, 320 px: ragged third tile
    (2, 96, 128, 10, 14, True),     # RAFT-small head, 140 px
    (1, 100, 64, 8, 16, True),      # last channel block partially out of range (100 = 3*32 + 4), exactly one tile
    (3, 32, 48, 9, 12, False),      # D padded 48 -> 64 (padding must read as zero), no bias, 108 px < one tile
    (1, 8, 16, 8, 8, True),         # smallest everything
]


@pytest.mark.parametrize("shape", SEEDED)
def test_operand_buffers_vs_oracle(pkg, shape):
    B, Cin, D, H, W, has_bias = shape
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2 * B, Cin, H, W)).astype(np.float32)
    w = (rng.normal(size=(D, Cin)) / np.sqrt(Cin)).astype(np.float32)
    b = rng.normal(size=(D,)).astype(np.float32) if has_bias else None
    f1, f2 = O.fnet_head(x, w, b, np.float64)
    L = 2 if min(H, W) < 16 else 4
    f1n, f2n, pyr = _operands(pkg, t(x), t(w), t(b) if has_bias else None, B, Cin, D, H, W, L)
    for fn, ref in ((f1n, f1), (f2n, f2)):
        got = fn.cpu().numpy()
        want = ref.reshape(B, D, H * W).transpose(0, 2, 1)
        assert np.abs(got[..., :D] - want).max() <= TOL_OPERAND * np.abs(want).max()
        assert not got[..., D:].any(), "K padding must be zero"
        assert not (got.view(np.uint32) & 0x1FFF).any(), "operands must be TF32-rounded"
    # and the pyramid the build made from them
    v0 = O.corr_volume(f1, f2, np.float64).reshape(B, H * W, H, W)
    blk = pkg.CorrBlock.from_encoder_head(t(x), t(w), t(b) if has_bias else None, num_levels=L, radius=2)
    assert rel_err(blk.corr_pyramid[0].reshape(B, H * W, H, W), v0) <= TOL
    # same call through the Python surface: same bits (compare the levels, the buffer's padding is never written)
    blk2 = pkg.CorrBlock.from_encoder_head(t(x), t(w), t(b) if has_bias else None, num_levels=L, radius=2)
    blk2._pyr = pyr
    for a, c in zip(blk.corr_pyramid, blk2.corr_pyramid):
        assert torch.equal(a, c)


def test_head_build_operand_modes_agree(pkg, monkeypatch):
    """After the head, the build runs on the head's TF32 operands (small maps) or on block-scaled fp16 copies of them
    (from 64x64 maps on; RAFTCORR_HEAD_F16 forces either).  TF32-rounded values times a power of two are exact in fp16,
    so the two volumes agree to the MMA's summation order."""
    B, Cin, D, H, W, L = 2, 96, 128, 24, 40, 3
    g = torch.Generator(device=dev()).manual_seed(51)
    x = torch.randn((2 * B, Cin, H, W), generator=g, device=dev()) * 37.0
    w = torch.randn((D, Cin), generator=g, device=dev()) / Cin ** 0.5
    pyr = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("RAFTCORR_HEAD_F16", flag)
        pyr[flag] = pkg.CorrBlock.from_encoder_head(x, w, None, num_levels=L, radius=3).corr_pyramid
    for a, c in zip(pyr["0"], pyr["1"]):
        assert float((a - c).abs().max()) <= 2e-6 * float(a.abs().max())


def test_unsupported_shape_goes_through_conv2(pkg):
    """H*W % 4 != 0: the fused kernel's TMA strides do not exist; the constructor computes conv2 itself (still on the GPU)."""
    B, Cin, D, H, W = 1, 32, 64, 9, 11
    from optical_flow_dnn_b200 import _lib

    assert not _lib.lib().raftcorr_fnet_head_supported(Cin, D, H, W)
    assert _lib.lib().raftcorr_fnet_head_supported(128, 256, 55, 128) and _lib.lib().raftcorr_fnet_head_supported(96, 128, 46, 62)
    rng = np.random.default_rng(6)
    x = rng.normal(size=(2 * B, Cin, H, W)).astype(np.float32)
    w = (rng.normal(size=(D, Cin, 1, 1)) / np.sqrt(Cin)).astype(np.float32)
    f1, f2 = O.fnet_head(x, w, None, np.float64)
    blk = pkg.CorrBlock.from_encoder_head(t(x), t(w), None, num_levels=2, radius=3)
    assert rel_err(blk.corr_pyramid[0].reshape(B, H * W, H, W), O.corr_volume(f1, f2, np.float64).reshape(B, H * W, H, W)) <= TOL


def test_full_size_against_unfused(pkg):
    """436x1024 -> 55x128, Cin=128 -> D=256, r=4 (RAFT-basic, BASELINE.json configs[1] shape), B=2 pairs."""
    B, Cin, D, H, W, L, r = 2, 128, 256, 55, 128, 4, 4
    g = torch.Generator(device=dev()).manual_seed(41)
    x = torch.randn((2 * B, Cin, H, W), generator=g, device=dev())
    w = torch.randn((D, Cin, 1, 1), generator=g, device=dev()) / Cin ** 0.5
    b = 0.1 * torch.randn((D,), generator=g, device=dev())
    coords = torch.from_numpy(O.coords_grid(B, H, W)).to(dev()) + 4.0 * torch.randn((B, 2, H, W), generator=g, device=dev())
    with torch.no_grad():
        f = torch.nn.functional.conv2d(x.double(), w.double(), b.double()).float()
        ref_blk = pkg.CorrBlock(f[:B], f[B:], L, r, build_mode="fp32")
        ref = ref_blk(coords)
        blk = pkg.CorrBlock.from_encoder_head(x, w, b, num_levels=L, radius=r)
        out = blk(coords)
        assert rel_err(out, ref) <= TOL
        # size-independent property: the centre tap at integer coordinates is the volume entry itself
        grid = torch.from_numpy(O.coords_grid(B, H, W)).to(dev())
        centre = blk(grid)[:, 40]  # level 0, i = j = r
        diag = torch.einsum("bdn,bdn->bn", f[:B].reshape(B, D, -1).double(), f[B:].reshape(B, D, -1).double()) / D ** 0.5
        assert rel_err(centre.reshape(B, -1), diag) <= TOL
        # fused lookup + convc1 runs on a head-built block as on any other
        wc = torch.randn((256, L * 81), generator=g, device=dev()) / 18.0
        assert rel_err(blk.lookup_convc1(coords, wc), torch.relu(torch.einsum("oc,bchw->bohw", wc.double(), out.double()))) <= 1e-3


def test_autograd_matches_unfused_fp32_path(pkg):
    B, Cin, D, H, W, L, r = 2, 96, 128, 16, 20, 4, 3
    g = torch.Generator(device=dev()).manual_seed(43)
    x = torch.randn((2 * B, Cin, H, W), generator=g, device=dev())
    w = torch.randn((D, Cin, 1, 1), generator=g, device=dev()) / Cin ** 0.5
    b = 0.1 * torch.randn((D,), generator=g, device=dev())
    coords = [torch.from_numpy(O.coords_grid(B, H, W)).to(dev()) + s * torch.randn((B, 2, H, W), generator=g, device=dev())
              for s in (0.0, 2.0, 4.0)]
    gout = [torch.randn((B, L * (2 * r + 1) ** 2, H, W), generator=g, device=dev()) for _ in coords]

    def run(fused):
        xs, ws, bs = (z.detach().clone().requires_grad_(True) for z in (x, w, b))
        if fused:
            blk = pkg.CorrBlock.from_encoder_head(xs, ws, bs, num_levels=L, radius=r)
        else:
            old = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            f = torch.nn.functional.conv2d(xs, ws, bs)
            torch.backends.cudnn.allow_tf32 = old
            blk = pkg.CorrBlock(f[:B], f[B:], L, r, build_mode="fp32")
        loss = sum((blk(c) * go).sum() for c, go in zip(coords, gout))
        loss.backward()
        return xs.grad, ws.grad, bs.grad

    got, want = run(True), run(False)
    for name, a, ref in zip(("dx", "dweight", "dbias"), got, want):
        assert a is not None and a.shape == ref.shape
        assert rel_err(a, ref) <= 2e-3, name   # TF32 forward volume is not involved; TF32 backward GEMMs + cuDNN's conv adjoint


def test_errors(pkg):
    d = dev()
    x = torch.randn(2, 32, 16, 16, device=d)
    with pytest.raises(RuntimeError):
        pkg.CorrBlock.from_encoder_head(x[:1], torch.randn(64, 32, device=d))        # odd image count
    with pytest.raises(RuntimeError):
        pkg.CorrBlock.from_encoder_head(x, torch.randn(64, 31, device=d))            # Cin mismatch
    with pytest.raises(RuntimeError):
        pkg.CorrBlock.from_encoder_head(x.cpu(), torch.randn(64, 32, device=d))      # CPU input: no CPU path
    with pytest.raises(RuntimeError):
        pkg.CorrBlock.from_encoder_head(x, torch.randn(64, 32, device=d), torch.randn(63, device=d))


def test_alternate_block_from_encoder_head(pkg, golden_fnet_head):
    """The on-the-fly class built from the head: same goldens (AlternateCorrBlock == CorrBlock)."""
    g = golden_fnet_head
    B, Cin, D, H, W, L, r = [int(v) for v in g["meta"]]
    with torch.no_grad():
        blk = pkg.AlternateCorrBlock.from_encoder_head(t(g["x"]), t(g["weight"]), t(g["bias"]), num_levels=L, radius=r)
        for k in range(3):
            out = blk(t(g[f"coords{k}"]))
            assert tuple(out.shape) == (B, L * (2 * r + 1) ** 2, H, W)
            assert rel_err(out, g[f"out64_{k}"]) <= TOL, f"coords set {k}"
        f1, f2 = blk.pyramid[0]
        assert rel_err(f1, g["fmap1"]) <= 1e-3 and rel_err(f2, g["fmap2"]) <= 1e-3


def test_alternate_block_head_autograd(pkg):
    B, Cin, D, H, W, L, r = 1, 96, 128, 16, 20, 3, 3
    g = torch.Generator(device=dev()).manual_seed(47)
    x = torch.randn((2 * B, Cin, H, W), generator=g, device=dev())
    w = torch.randn((D, Cin, 1, 1), generator=g, device=dev()) / Cin ** 0.5
    b = 0.1 * torch.randn((D,), generator=g, device=dev())
    coords = [torch.from_numpy(O.coords_grid(B, H, W)).to(dev()) + s * torch.randn((B, 2, H, W), generator=g, device=dev())
              for s in (0.0, 3.0)]
    gout = [torch.randn((B, L * (2 * r + 1) ** 2, H, W), generator=g, device=dev()) for _ in coords]

    def run(fused):
        xs, ws, bs = (z.detach().clone().requires_grad_(True) for z in (x, w, b))
        if fused:
            blk = pkg.AlternateCorrBlock.from_encoder_head(xs, ws, bs, num_levels=L, radius=r)
        else:
            old = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            f = torch.nn.functional.conv2d(xs, ws, bs)
            torch.backends.cudnn.allow_tf32 = old
            blk = pkg.CorrBlock(f[:B], f[B:], L, r, build_mode="fp32")
        sum((blk(c) * go).sum() for c, go in zip(coords, gout)).backward()
        return xs.grad, ws.grad, bs.grad

    got, want = run(True), run(False)
    for name, a, ref in zip(("dx", "dweight", "dbias"), got, want):
        assert a is not None and a.shape == ref.shape
        assert rel_err(a, ref) <= 2e-3, name
