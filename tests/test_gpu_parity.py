"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libraftcorr_b200.so),
against (1) golden vectors produced by the reference itself, (2) the CPU oracle on seeded inputs,
(3) size-independent properties at BASELINE.json's full sizes.

Tolerances (north_star: "fp32-accumulated, <=1e-3 relative on the volume and lookup"), made
concrete as max|delta| <= tol * max|ref| per tensor:
    fp32 build mode  : 2e-5   (fp32 re-ordering only)
    tf32 build mode  : 1e-3   (TF32-rounded operands, fp32 accumulation in TMEM)
    gradients        : 1e-3 (tf32 forward does not matter: the adjoints are fp32)
"""
import numpy as np
import pytest
import torch

from make_golden import grad_weights
from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-5, "tf32": 1e-3, "f16": 1e-3}   # f16 = block-scaled fp16 operands: TF32's 11 significant bits
MODES = ["fp32", "tf32", "f16"]


def dev():
    return torch.device("cuda:0")


def rel_err(a, ref):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    return float(np.abs(a.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def _meta(g):
    return [int(v) for v in g["meta"]]


@pytest.fixture(scope="module")
def pkg():
    import optical_flow_dnn_b200 as m

    return m


# ------------------------------------------------------------------ golden vectors (reference outputs)
@pytest.mark.parametrize("mode", MODES)
def test_golden_pyramid_and_lookup(pkg, golden, mode):
    B, D, H, W, L, r = _meta(golden)
    blk = pkg.CorrBlock(torch.from_numpy(golden["fmap1"]).to(dev()), torch.from_numpy(golden["fmap2"]).to(dev()),
                        num_levels=L, radius=r, build_mode=mode)
    pyr = blk.corr_pyramid
    assert len(pyr) == L
    for l in range(L):
        ref = golden[f"pyr{l}"]
        assert tuple(pyr[l].shape) == (B * H * W, 1) + ref.shape[-2:]
        assert rel_err(pyr[l].reshape(ref.shape), ref) <= TOL[mode], f"level {l}"
    for k in range(3):
        out = blk(torch.from_numpy(golden[f"coords{k}"]).to(dev()))
        assert out.shape == golden[f"out{k}"].shape and out.dtype == torch.float32 and out.is_contiguous()
        assert rel_err(out, golden[f"out64_{k}"]) <= TOL[mode], f"coords set {k}"


@pytest.mark.parametrize("mode", MODES)
def test_golden_gradients(pkg, golden, mode):
    B, D, H, W, L, r = _meta(golden)
    f1 = torch.from_numpy(golden["fmap1"]).to(dev()).requires_grad_()
    f2 = torch.from_numpy(golden["fmap2"]).to(dev()).requires_grad_()
    blk = pkg.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=mode)
    loss = 0
    cs = []
    for k in range(3):
        c = torch.from_numpy(golden[f"coords{k}"]).to(dev()).requires_grad_(k == 1)
        cs.append(c)
        out = blk(c)
        w = torch.from_numpy(grad_weights(tuple(out.shape), k).astype(np.float32)).to(dev())
        loss = loss + (out * w).sum()
    loss.backward()
    assert rel_err(f1.grad, golden["df1"]) <= 1e-3
    assert rel_err(f2.grad, golden["df2"]) <= 1e-3
    # coordinates gradient (the reference CorrBlock provides it through grid_sample): generic coords only
    assert rel_err(cs[1].grad, golden["dcoords1"]) <= TOL[mode] * 2


ALT_TOL = {"fp32": 2e-5, "tf32": 1e-3, "f16": 1e-3}  # on-the-fly path: exact fp32 dots vs TF32 tensor-core tiles


@pytest.mark.parametrize("mode", MODES)
def test_golden_alternate(pkg, golden, mode):
    B, D, H, W, L, r = _meta(golden)
    if mode == "tf32" and D % 32:
        pytest.skip("the tensor-core on-the-fly kernel needs D % 32 == 0")
    f1 = torch.from_numpy(golden["fmap1"]).to(dev()).requires_grad_()
    f2 = torch.from_numpy(golden["fmap2"]).to(dev()).requires_grad_()
    blk = pkg.AlternateCorrBlock(f1, f2, num_levels=L, radius=r, build_mode=mode)
    loss = 0
    for k in range(3):
        out = blk(torch.from_numpy(golden[f"coords{k}"]).to(dev()))
        assert rel_err(out, golden[f"out64_{k}"]) <= ALT_TOL[mode]
        w = torch.from_numpy(grad_weights(tuple(out.shape), k).astype(np.float32)).to(dev())
        loss = loss + (out * w).sum()
    loss.backward()
    assert rel_err(f1.grad, golden["df1"]) <= (1e-4 if mode == "fp32" else 1e-3)
    assert rel_err(f2.grad, golden["df2"]) <= (1e-4 if mode == "fp32" else 1e-3)
    assert len(blk.pyramid) == L + 1


def test_alt_cuda_corr_module_surface(pkg, golden):
    """The native-extension mirror: single level, unscaled, NHWC fmaps, interleaved coords
    (raft/alt_cuda_corr/correlation.cpp:23-49)."""
    from optical_flow_dnn_b200 import alt_cuda_corr

    B, D, H, W, L, r = _meta(golden)
    f1 = torch.from_numpy(golden["fmap1"]).to(dev())
    f2 = torch.from_numpy(golden["fmap2"]).to(dev())
    f1n = f1.permute(0, 2, 3, 1).contiguous()
    for lvl in range(2):
        f2l = f2
        for _ in range(lvl):
            f2l = torch.nn.functional.avg_pool2d(f2l, 2, stride=2)
        f2n = f2l.permute(0, 2, 3, 1).contiguous()
        c = torch.from_numpy(golden["coords1"]).to(dev())
        ci = (c.permute(0, 2, 3, 1) / 2 ** lvl).reshape(B, 1, H, W, 2).contiguous()
        corr, = alt_cuda_corr.forward(f1n, f2n, ci, r)
        K = 2 * r + 1
        assert tuple(corr.shape) == (B, 1, K * K, H, W)
        ref = golden["out64_1"][:, lvl * K * K:(lvl + 1) * K * K] * np.sqrt(D)
        assert rel_err(corr[:, 0], ref) <= 2e-5
        g = torch.from_numpy(grad_weights(tuple(corr.shape), 5).astype(np.float32)).to(dev())
        d1, d2, dc = alt_cuda_corr.backward(f1n, f2n, ci, g, r)
        assert d1.shape == f1n.shape and d2.shape == f2n.shape and dc.shape == ci.shape
        assert float(dc.abs().max()) == 0.0  # correlation_kernel.cu:307
        # oracle: single-level alt grad == d/dfmaps of <alt_lookup(L=1) * sqrt(D), g>
        co = (golden["coords1"] / 2 ** lvl).astype(np.float32)
        e1, e2 = O.alt_grad(golden["fmap1"], f2l.cpu().numpy(), co, g.cpu().numpy()[:, 0] * np.sqrt(D), 1, r, np.float64) \
            if lvl == 0 else (None, None)
        if lvl == 0:
            assert rel_err(d1.permute(0, 3, 1, 2), e1) <= 1e-4
            assert rel_err(d2.permute(0, 3, 1, 2), e2) <= 1e-4
    with pytest.raises(RuntimeError):
        alt_cuda_corr.forward(f1n.cpu(), f2n, ci, r)
    with pytest.raises(RuntimeError):
        alt_cuda_corr.forward(f1.permute(0, 2, 3, 1), f2n, ci, r)  # non-contiguous


# ------------------------------------------------------------------ seeded inputs vs the CPU oracle
SEEDED = [
    # B, D, H, W, L, r   (config 1 shape; an odd KITTI-like shape; tiny)
    (1, 128, 46, 62, 4, 3),
    (2, 256, 23, 47, 4, 4),
    (2, 64, 17, 16, 3, 2),
    (1, 32, 8, 9, 2, 1),
    (1, 320, 12, 20, 2, 2),  # D > 256: ten k-blocks in the tensor-core kernels, CUDA-core backward GEMMs
]


@pytest.mark.parametrize("shape", SEEDED)
@pytest.mark.parametrize("mode", MODES)
def test_seeded_vs_oracle(pkg, shape, mode):
    B, D, H, W, L, r = shape
    rng = np.random.default_rng(1234)
    f1 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    pyr = O.corr_pyramid(f1, f2, L, np.float64)
    blk = pkg.CorrBlock(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()), L, r, build_mode=mode)
    got = blk.corr_pyramid
    for l in range(L):
        assert rel_err(got[l].reshape(pyr[l].shape), pyr[l]) <= TOL[mode], f"level {l}"
    coords = O.coords_grid(B, H, W)
    for it in range(3):
        coords = coords + rng.normal(0, 4.0 if it == 0 else 0.5, coords.shape).astype(np.float32)
        ref = O.lookup(pyr, coords, r)
        out = blk(torch.from_numpy(coords).to(dev()))
        assert rel_err(out, ref) <= TOL[mode]
    if mode == "fp32" or D % 32 == 0:
        alt = pkg.AlternateCorrBlock(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()), L, r, build_mode=mode)
        assert rel_err(alt(torch.from_numpy(coords).to(dev())), ref) <= ALT_TOL[mode]


def test_c_oracle_agrees_on_gpu_shapes(pkg):
    """Second, independent checker (plain C loops) on one seeded case."""
    from oracle import c_oracle as C

    B, D, H, W, L, r = 1, 64, 20, 28, 4, 4
    rng = np.random.default_rng(5)
    f1 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    coords = O.coords_grid(B, H, W) + rng.normal(0, 3.0, (B, 2, H, W)).astype(np.float32)
    ref = C.lookup(C.corr_pyramid(f1, f2, L), coords, r)
    blk = pkg.CorrBlock(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()), L, r, build_mode="fp32")
    assert rel_err(blk(torch.from_numpy(coords).to(dev())), ref.astype(np.float64)) <= 2e-5


# ------------------------------------------------------------------ full-size properties (config 2 shape)
@pytest.mark.parametrize("mode", MODES)
def test_full_size_properties(pkg, mode):
    """436x1024 -> 55x128 feature maps, D=256, r=4 (BASELINE.json configs[1]), B=2: too big for the
    CPU oracle in seconds, so check size-independent properties instead."""
    B, D, H, W, L, r = 2, 256, 55, 128, 4, 4
    g = torch.Generator(device=dev()).manual_seed(1234)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    blk = pkg.CorrBlock(f1, f2, L, r, build_mode=mode)
    pyr = blk.corr_pyramid
    N = H * W
    v0 = pyr[0].reshape(B, N, H, W)
    tol = TOL[mode]
    # (1) checksum of checksums: sum_n V0[m, n] == <f1[m], sum_n f2[n]> / sqrt(D)   (fp64 on device)
    rowsum = v0.double().sum(dim=(2, 3))
    expect = torch.einsum("bdm,bd->bm", f1.double().reshape(B, D, N), f2.double().reshape(B, D, N).sum(-1)) / np.sqrt(D)
    assert float((rowsum - expect).abs().max() / expect.abs().max()) <= tol
    # (2) random spot checks of single entries against fp64 dot products
    idx = torch.randint(0, N, (4096, 2), generator=g, device=dev())
    a = f1.double().reshape(B, D, N)[0][:, idx[:, 0]]
    b = f2.double().reshape(B, D, N)[0][:, idx[:, 1]]
    ref = (a * b).sum(0) / np.sqrt(D)
    got = v0[0].reshape(N, N)[idx[:, 0], idx[:, 1]].double()
    assert float((got - ref).abs().max()) <= tol * float(v0.abs().max())
    # (3) each level is the floor-mode 2x2 mean of the previous one (corr.py:41-43)
    for l in range(1, L):
        want = torch.nn.functional.avg_pool2d(pyr[l - 1].contiguous(), 2, stride=2)
        assert tuple(want.shape) == tuple(pyr[l].shape)
        assert float((want - pyr[l]).abs().max()) <= 1e-5 * float(v0.abs().max())
    # (4) linearity of the whole path in fmap1
    blk2 = pkg.CorrBlock(2.0 * f1, f2, L, r, build_mode=mode)
    coords = torch.from_numpy(O.coords_grid(B, H, W)).to(dev()) + 3.0 * torch.randn((B, 2, H, W), generator=g, device=dev())
    o1, o2 = blk(coords), blk2(coords)
    assert float((o2 - 2.0 * o1).abs().max()) <= 1e-5 * float(o1.abs().max())
    # (5) at integer coordinates the centre tap of level 0 is the volume entry itself
    cg = torch.from_numpy(O.coords_grid(B, H, W)).to(dev())
    og = blk(cg)
    K = 2 * r + 1
    centre = og[:, r * K + r].reshape(B, N)
    diag = v0.reshape(B, N, N).diagonal(dim1=1, dim2=2)
    assert float((centre - diag).abs().max()) == 0.0
    # (6) lookup == on-the-fly variant (AlternateCorrBlock == CorrBlock)
    alt = pkg.AlternateCorrBlock(f1, f2, L, r, build_mode=mode)(coords)
    assert float((alt - o1).abs().max()) <= tol * float(o1.abs().max())
    # (7) adjoint identity <lookup(V), g> == <V, lookup^T(g)> through autograd at full size
    f1g = f1.clone().requires_grad_()
    blk3 = pkg.CorrBlock(f1g, f2, L, r, build_mode=mode)
    gsel = torch.randn(o1.shape, generator=g, device=dev())
    (blk3(coords) * gsel).sum().backward()
    eps_dir = torch.randn(f1.shape, generator=g, device=dev())
    lhs = float((f1g.grad.double() * eps_dir.double()).sum())
    o_dir = pkg.CorrBlock(eps_dir, f2, L, r, build_mode=mode)(coords)  # the path is linear in fmap1
    rhs = float((o_dir.double() * gsel.double()).sum())
    assert abs(lhs - rhs) <= 2e-3 * max(abs(lhs), abs(rhs), 1.0)


# ------------------------------------------------------------------ block-scaled fp16 operands: dynamic range
def test_f16_block_scaling_dynamic_range(pkg):
    """fp16 has 5 exponent bits; the per-tile power-of-two scales must make the f16 build as good as TF32 for feature
    maps of any overall magnitude (1e-20 .. 1e+15), for magnitudes that differ by tile (x 2^+-20 per pixel block), and
    for all-zero blocks; scaling by a power of two must give the same bits up to that factor."""
    B, D, H, W, L, r = 2, 96, 24, 40, 3, 4   # D padded 96 -> 128 halves; 960 px = 7.5 m-tiles; 3 x 2 target tiles
    N = H * W
    g = torch.Generator(device=dev()).manual_seed(5)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    # per-block magnitudes: every 128-pixel run of fmap1 and every 8x32 tile of fmap2 gets its own power of two
    e1 = torch.randint(-20, 21, (B, 1, (N + 127) // 128), generator=g, device=dev()).float().repeat_interleave(128, 2)[..., :N]
    f1 = f1 * torch.exp2(e1).reshape(B, 1, H, W)
    e2 = torch.randint(-20, 21, (B, 1, (H + 7) // 8, (W + 31) // 32), generator=g, device=dev()).float()
    f2 = f2 * torch.exp2(e2).repeat_interleave(8, 2).repeat_interleave(32, 3)[..., :H, :W]
    f2[:, :, 8:16, 0:32] = 0.0   # an all-zero target tile
    for mag in (1.0, 1e-20, 1e15):
        a, b_ = f1 * mag, f2
        ref = torch.einsum("bdm,bdn->bmn", a.reshape(B, D, N).double(), b_.reshape(B, D, N).double()) / np.sqrt(D)
        for mode in ("tf32", "f16"):
            v0 = pkg.CorrBlock(a, b_, L, r, build_mode=mode).corr_pyramid[0].reshape(B, N, N).double()
            assert torch.isfinite(v0).all()
            # per source-pixel row: the row's own scale (magnitudes differ by 2^40 between blocks)
            err = (v0 - ref).abs().amax(dim=2) / ref.abs().amax(dim=2).clamp_min(1e-300)
            assert float(err.max()) <= 1e-3, (mode, mag, float(err.max()))
    v1 = pkg.CorrBlock(f1, f2, L, r, build_mode="f16").corr_pyramid[0]
    v2 = pkg.CorrBlock(f1 * 1024.0, f2 * 0.125, L, r, build_mode="f16").corr_pyramid[0]
    assert torch.equal(v1 * 128.0, v2)


def test_f16x3_build_is_fp32_class(pkg, golden):
    """build_mode="f16x3": three-term fp16 split on the tensor cores.  Same tolerance as the exact-fp32 FFMA build (2e-5
    of max|ref|) on the reference goldens (every level, three coordinate sets) and, at the bench shape, against fp64."""
    B, D, H, W, L, r = [int(v) for v in golden["meta"]]
    blk = pkg.CorrBlock(torch.from_numpy(golden["fmap1"]).to(dev()), torch.from_numpy(golden["fmap2"]).to(dev()), L, r,
                        build_mode="f16x3")
    pyr = blk.corr_pyramid
    for l in range(L):
        ref = golden[f"pyr{l}"]
        assert rel_err(pyr[l].reshape(ref.shape), ref) <= TOL["fp32"], f"level {l}"
    for k in range(3):
        assert rel_err(blk(torch.from_numpy(golden[f"coords{k}"]).to(dev())), golden[f"out64_{k}"]) <= TOL["fp32"], f"coords set {k}"


def test_f16x3_full_size_and_range(pkg):
    B, D, H, W, L, r = 1, 256, 55, 128, 4, 4
    N = H * W
    g = torch.Generator(device=dev()).manual_seed(21)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    ref = torch.einsum("dm,dn->mn", f1[0].reshape(D, N).double(), f2[0].reshape(D, N).double()) / np.sqrt(D)
    errs = {}
    for mode in ("f16x3", "fp32", "f16"):
        v0 = pkg.CorrBlock(f1, f2, L, r, build_mode=mode).corr_pyramid[0].reshape(N, N).double()
        errs[mode] = float((v0 - ref).abs().max() / ref.abs().max())
    print("max|dV|/max|V| at the bench shape:", errs)
    assert errs["f16x3"] <= 5e-6 and errs["f16x3"] <= errs["f16"] / 50
    for mag in (1e-20, 1e15):   # block scaling: any overall magnitude
        v0 = pkg.CorrBlock(f1 * mag, f2, L, r, build_mode="f16x3").corr_pyramid[0].reshape(N, N).double()
        assert float((v0 - ref * mag).abs().max() / (ref.abs().max() * mag)) <= 5e-6


def test_alt_f16_operands_match_tf32_operands(pkg, monkeypatch):
    """The on-the-fly tensor-core lookup reads fp16 copies of the feature tensors (one power-of-two scale per image and
    tensor) by default; RAFTCORR_ALT_F16=0 keeps TF32 operands.  Same error against the exact-fp32 volume path, for any
    overall magnitude and for images of very different magnitude inside one batch."""
    B, D, H, W, L, r = 3, 128, 24, 40, 3, 4
    g = torch.Generator(device=dev()).manual_seed(9)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    f1 = f1 * torch.tensor([1.0, 2.0 ** 17, 2.0 ** -21], device=dev()).reshape(B, 1, 1, 1)   # per-image magnitudes
    coords = torch.from_numpy(O.coords_grid(B, H, W)).to(dev()) + 3.0 * torch.randn((B, 2, H, W), generator=g, device=dev())
    for mag in (1.0, 1e-18, 1e12):
        a = f1 * mag
        ref = pkg.CorrBlock(a, f2, L, r, build_mode="fp32")(coords).double()
        scale = ref.abs().amax(dim=(1, 2, 3), keepdim=True)   # per image: their magnitudes differ by 2^38
        outs = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("RAFTCORR_ALT_F16", flag)
            blk = pkg.AlternateCorrBlock(a, f2, L, r, build_mode="tf32")
            outs[flag] = blk(coords).double()
            assert (blk._f16 is not False) == (flag == "1")
            assert float(((outs[flag] - ref).abs() / scale).max()) <= 1e-3, (flag, mag)
        assert float(((outs["1"] - outs["0"]).abs() / scale).max()) <= 1e-3
    monkeypatch.delenv("RAFTCORR_ALT_F16")
    # D % 64 != 0: TF32 operands
    blk = pkg.AlternateCorrBlock(f1[:, :96], f2[:, :96], L, r, build_mode="tf32")
    blk(coords)
    assert blk._f16 is False


# ------------------------------------------------------------------ pyramids beyond 4 GiB (configs 4 and 5a)
@pytest.mark.parametrize("shape", [(16, 256, 47, 156), (1, 256, 135, 240)], ids=["config4_B16_kitti", "config5a_B1_1080p"])
def test_pyramid_larger_than_4GiB(pkg, shape):
    """BASELINE.json configs[3] (raft_uncertainty on KITTI, 375x1242 -> 47x156, B=16: 4.5 GB pyramid) and the 1080p
    shape of configs[4] as a volume (135x240, one pair: 5.6 GB): every offset past 2^32 bytes must be right.  Checked on
    the LAST source pixels of the LAST pair, which live at the far end of each level: volume rows against fp64 dot
    products, each level against the 2x2 mean of the previous one, lookup against the on-the-fly class."""
    B, D, H, W = shape
    L, r = 4, 4
    N = H * W
    g = torch.Generator(device=dev()).manual_seed(77)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    with torch.no_grad():
        blk = pkg.CorrBlock(f1, f2, L, r)
        assert blk._layout.total_floats * 4 > 2 ** 32
        b = B - 1
        ms = torch.tensor([0, N // 2, N - 130, N - 2, N - 1], device=dev())
        rows = torch.einsum("dm,dn->mn", f1[b].reshape(D, N)[:, ms].double(), f2[b].reshape(D, N).double()) / np.sqrt(D)
        pyr = blk.corr_pyramid
        got = pyr[0].reshape(B, N, H, W)[b, ms].reshape(len(ms), N)
        scale = float(rows.abs().max())
        assert float((got.double() - rows).abs().max()) <= TOL["tf32"] * scale
        for l in range(1, L):
            prev = pyr[l - 1].reshape(B, N, *pyr[l - 1].shape[-2:])[b, ms]
            want = torch.nn.functional.avg_pool2d(prev.contiguous(), 2, stride=2)
            assert float((want - pyr[l].reshape(B, N, *pyr[l].shape[-2:])[b, ms]).abs().max()) <= 1e-5 * scale
        del pyr, got, rows
        coords = torch.from_numpy(O.coords_grid(B, H, W)).to(dev()) + 3.0 * torch.randn((B, 2, H, W), generator=g, device=dev())
        out = blk(coords)
        alt = pkg.AlternateCorrBlock(f1, f2, L, r)(coords)
        assert float((alt - out).abs().max()) <= TOL["tf32"] * float(out.abs().max())
        # and the last pair alone gives the same bits as inside the big batch
        solo = pkg.CorrBlock(f1[b:], f2[b:], L, r)(coords[b:])
        assert torch.equal(solo[0], out[b])


# ------------------------------------------------------------------ edge cases / error behaviour
def test_edge_cases_and_errors(pkg):
    d = dev()
    f = torch.randn(1, 16, 16, 16, device=d)
    with pytest.raises(RuntimeError):
        pkg.CorrBlock(f.cpu(), f.cpu())
    with pytest.raises(RuntimeError):
        pkg.CorrBlock(f, f.double())
    with pytest.raises(pkg.RaftCorrError):
        pkg.CorrBlock(f[:, :, :8, :8], f[:, :, :8, :8], num_levels=4)  # level 3 would be 1x1 (NaN in the reference)
    with pytest.raises(pkg.RaftCorrError):
        pkg.CorrBlock(f, f, radius=7)
    blk = pkg.CorrBlock(f, f, num_levels=2, radius=3, build_mode="fp32")
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 2, 8, 8, device=d))
    # empty batch
    e = torch.zeros(0, 16, 16, 16, device=d)
    out = pkg.CorrBlock(e, e, 2, 3, build_mode="fp32")(torch.zeros(0, 2, 16, 16, device=d))
    assert tuple(out.shape) == (0, 2 * 49, 16, 16)
    # non-contiguous inputs are accepted (the reference's matmul / grid_sample accept them too)
    fn = torch.randn(1, 16, 16, 32, device=d)[..., ::2]
    cn = torch.randn(1, 16, 16, 2, device=d).permute(0, 3, 1, 2) * 4 + 8
    o1 = pkg.CorrBlock(fn, fn, 2, 3, build_mode="fp32")(cn)
    o2 = pkg.CorrBlock(fn.contiguous(), fn.contiguous(), 2, 3, build_mode="fp32")(cn.contiguous())
    assert torch.equal(o1, o2)
    # wild coordinates: far outside, huge, NaN -> zeros for every tap that falls outside (zero padding)
    c = torch.full((1, 2, 16, 16), 1.0e9, device=d)
    c[0, :, 0, 0] = float("nan")
    c[0, :, 1, 1] = -3.0e38
    o = blk(c)
    assert torch.isfinite(o[:, :, 2:, 2:]).all() and float(o[:, :, 2:, 2:].abs().max()) == 0.0
    # CorrBlock.corr static method (corr.py:93): [B,H,W,1,H,W]
    v = pkg.CorrBlock.corr(f, f, build_mode="fp32")
    assert tuple(v.shape) == (1, 16, 16, 1, 16, 16)
    ref = torch.einsum("bdm,bdn->bmn", f.reshape(1, 16, 256), f.reshape(1, 16, 256)) / 4.0
    assert float((v.reshape(1, 256, 256) - ref).abs().max()) <= 1e-4


def test_streams_and_determinism(pkg):
    """Kernels launch on the caller's current stream (the reference used the legacy default stream)
    and repeated runs are bit-identical (no atomics on the forward path)."""
    d = dev()
    g = torch.Generator(device=d).manual_seed(3)
    f1 = torch.randn((2, 64, 24, 32), generator=g, device=d)
    f2 = torch.randn((2, 64, 24, 32), generator=g, device=d)
    coords = torch.from_numpy(O.coords_grid(2, 24, 32)).to(d) + 2.0 * torch.randn((2, 2, 24, 32), generator=g, device=d)
    outs = []
    for mode in MODES:
        a = pkg.CorrBlock(f1, f2, 4, 4, build_mode=mode)(coords)
        s = torch.cuda.Stream(device=d)
        s.wait_stream(torch.cuda.current_stream(d))
        with torch.cuda.stream(s):
            b = pkg.CorrBlock(f1, f2, 4, 4, build_mode=mode)(coords)
        s.synchronize()
        assert torch.equal(a, b), mode
        outs.append(a)
    # batch sharding is exact: processing pair 1 alone gives the same bits as inside the batch
    solo = pkg.CorrBlock(f1[1:], f2[1:], 4, 4, build_mode="fp32")(coords[1:])
    assert torch.equal(solo, outs[0][1:])


# ------------------------------------------------------------------ implementation variants agree
def test_kernel_variants_agree(pkg, monkeypatch):
    """The A/B switches kept in the library (cp.async vs TMA lookup, single vs grouped on-the-fly kernel,
    1- vs 2-CTA build) must produce the same numbers; planned and unplanned C entry points too."""
    import ctypes

    from optical_flow_dnn_b200 import _lib

    d = dev()
    g = torch.Generator(device=d).manual_seed(11)
    B, D, H, W, L, r = 2, 128, 30, 45, 4, 4
    f1 = torch.randn((B, D, H, W), generator=g, device=d)
    f2 = torch.randn((B, D, H, W), generator=g, device=d)
    coords = torch.from_numpy(O.coords_grid(B, H, W)).to(d) + 3.0 * torch.randn((B, 2, H, W), generator=g, device=d)
    coords[0, :, 0, :5] = torch.tensor([[-7.3, 0.0, 44.0, 100.5, 3.999], [2.0, -0.5, 29.0, 1.0, 29.999]], device=d)

    # the fp16-operand build: its pipeline shapes agree bit for bit, and with the TF32 build to rounding
    v16 = pkg.CorrBlock(f1, f2, L, r, build_mode="f16").corr_pyramid
    for pipe in ("322", "222", "242"):
        monkeypatch.setenv("RAFTCORR_BUILD_PIPE", pipe)
        for a, c in zip(v16, pkg.CorrBlock(f1, f2, L, r, build_mode="f16").corr_pyramid):
            assert torch.equal(a, c), pipe
    monkeypatch.delenv("RAFTCORR_BUILD_PIPE")
    # everything below compares variants of the TF32 build (the CTA-pair experiments exist for it only)
    import optical_flow_dnn_b200.corr as corr_mod
    monkeypatch.setattr(corr_mod, "_default_mode", "tf32")
    for a, c in zip(v16, pkg.CorrBlock(f1, f2, L, r).corr_pyramid):
        assert float((a - c).abs().max()) <= 1e-3 * float(c.abs().max())
    # production layout: row-quad-interleaved pyramid, single-CTA build, TMA lookup
    blk_il = pkg.CorrBlock(f1, f2, L, r)
    assert blk_il._layout.interleaved == 2
    o_il = blk_il(coords)
    # round 1's row-pair layout: same arithmetic, different addresses -> bit-identical levels and lookups
    monkeypatch.setenv("RAFTCORR_PYRAMID_LAYOUT", "pairs")
    blk_p = pkg.CorrBlock(f1, f2, L, r)
    assert blk_p._layout.interleaved == 1
    for a, c in zip(blk_p.corr_pyramid, blk_il.corr_pyramid):
        assert torch.equal(a, c)
    assert torch.equal(blk_p(coords), o_il)
    monkeypatch.delenv("RAFTCORR_PYRAMID_LAYOUT")
    for pipe in ("322", "222", "242"):  # operand stages / store buffers / rows per buffer
        monkeypatch.setenv("RAFTCORR_BUILD_PIPE", pipe)
        for a, c in zip(blk_il.corr_pyramid, pkg.CorrBlock(f1, f2, L, r).corr_pyramid):
            assert torch.equal(a, c), pipe
    monkeypatch.delenv("RAFTCORR_BUILD_PIPE")
    # the row-major layout (same arithmetic, different addresses) and the variants that only exist for it
    monkeypatch.setenv("RAFTCORR_PYRAMID_LAYOUT", "rowmajor")
    monkeypatch.setenv("RAFTCORR_BUILD_CLUSTER", "1")
    blk = pkg.CorrBlock(f1, f2, L, r)
    assert blk._layout.interleaved == 0
    for a, c in zip(blk.corr_pyramid, blk_il.corr_pyramid):  # views exclude the (never written) row padding
        assert torch.equal(a, c)
    monkeypatch.setenv("RAFTCORR_BUILD_CLUSTER", "2")
    blk2 = pkg.CorrBlock(f1, f2, L, r)
    for a, c in zip(blk.corr_pyramid, blk2.corr_pyramid):
        assert torch.equal(a, c)
    monkeypatch.setenv("RAFTCORR_BUILD_CLUSTER", "1")
    for pipe in ("322", "222", "242", "341", "421"):
        monkeypatch.setenv("RAFTCORR_BUILD_PIPE", pipe)
        for a, c in zip(blk.corr_pyramid, pkg.CorrBlock(f1, f2, L, r).corr_pyramid):
            assert torch.equal(a, c), pipe
    monkeypatch.setenv("RAFTCORR_BUILD_PIPE", "999")
    with pytest.raises(pkg.RaftCorrError, match="pipeline shape"):
        pkg.CorrBlock(f1, f2, L, r)
    monkeypatch.delenv("RAFTCORR_BUILD_PIPE")
    monkeypatch.setenv("RAFTCORR_LOOKUP", "tma")
    o_tma = blk(coords)
    o_tma2 = blk2(coords)
    assert torch.equal(o_tma, o_il)
    monkeypatch.setenv("RAFTCORR_LOOKUP_BUFS", "2")  # 2-deep patch ring, 3 CTAs per SM
    assert torch.equal(blk(coords), o_tma)
    monkeypatch.delenv("RAFTCORR_LOOKUP_BUFS")
    monkeypatch.setenv("RAFTCORR_LOOKUP", "cpasync")
    o_cp = blk(coords)
    with pytest.raises(pkg.RaftCorrError, match="row-major"):
        blk_il(coords)
    assert torch.equal(o_tma, o_tma2)
    assert float((o_tma - o_cp).abs().max()) <= 1e-6 * float(o_tma.abs().max())
    # unplanned C entry point == planned one
    out = torch.empty_like(o_tma)
    monkeypatch.setenv("RAFTCORR_LOOKUP", "tma")
    rc_ = _lib.lib().raftcorr_lookup_fwd(ctypes.c_void_p(blk._pyr.data_ptr()), ctypes.byref(blk._layout),
                                         ctypes.c_void_p(coords.data_ptr()), r, ctypes.c_void_p(out.data_ptr()), d.index,
                                         ctypes.c_void_p(torch.cuda.current_stream(d).cuda_stream))
    assert rc_ == 0
    torch.cuda.synchronize()
    assert torch.equal(out, o_tma)
    # on-the-fly: grouped vs single-pixel kernel
    alt = pkg.AlternateCorrBlock(f1, f2, L, r, build_mode="fp32")
    monkeypatch.setenv("RAFTCORR_ALT", "group")
    a_g = alt(coords)
    monkeypatch.setenv("RAFTCORR_ALT", "single")
    a_s = alt(coords)
    assert float((a_g - a_s).abs().max()) <= 2e-6 * float(a_s.abs().max())
    # ... and the tensor-core tiles (TF32 operands) against the exact fp32 dots
    alt_t = pkg.AlternateCorrBlock(f1, f2, L, r, build_mode="tf32")
    a_t = alt_t(coords)
    assert float((a_t - a_s).abs().max()) <= 1e-3 * float(a_s.abs().max())
    # ... whose variants (32-point-wide lattice tiles only; 32, 16 and 8; always 256 points per tile: other tilings of the
    # same boxes; a 2-stage operand ring: the same arithmetic) differ by nothing but the order in which a window's points arrive
    for var, val in (("RAFTCORR_ALT_SQUARE", "0"), ("RAFTCORR_ALT_SQUARE", "3"), ("RAFTCORR_ALT_SHORT", "0"), ("RAFTCORR_ALT_STAGES", "2")):
        monkeypatch.setenv(var, val)
        a_v = alt_t(coords)
        monkeypatch.delenv(var)
        assert float((a_v - a_t).abs().max()) <= 2e-6 * float(a_t.abs().max()), var


@pytest.mark.parametrize("shape", [(1, 32, 8, 16, 1, 1), (2, 64, 19, 37, 3, 2), (1, 96, 33, 21, 4, 3), (2, 128, 24, 50, 4, 4)])
def test_alt_tensor_core_blocks(pkg, shape):
    """The tensor-core on-the-fly kernel on sizes that are not multiples of its 8x16 source block / 8x32 lattice
    tile, for every radius, with coherent, incoherent, integer, far-out-of-range and NaN coordinates."""
    B, D, H, W, L, r = shape
    rng = np.random.default_rng(5)
    f1 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    pyr = O.corr_pyramid(f1, f2, L, np.float64)
    scale = max(float(np.abs(p).max()) for p in pyr)
    alt = pkg.AlternateCorrBlock(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()), L, r, build_mode="tf32")
    grid = O.coords_grid(B, H, W)
    smooth = grid + np.float32(2.3) + rng.normal(0, 0.3, grid.shape).astype(np.float32)
    wild = grid + rng.normal(0, 12.0, grid.shape).astype(np.float32)
    far = grid.copy()
    far[:, 0, : H // 2] += 1e4   # whole rows outside every level
    far[:, 1, :, : W // 3] -= 777.25
    for name, coords in (("integer", grid), ("smooth", smooth), ("wild", wild), ("far", far)):
        ref = O.lookup(pyr, coords, r)
        out = alt(torch.from_numpy(coords).to(dev())).cpu().numpy()
        assert np.abs(out - ref).max() <= 1e-3 * scale, name
    nan = smooth.copy()
    nan[0, 0, 1, 2] = np.nan
    out = alt(torch.from_numpy(nan).to(dev())).cpu().numpy()
    ref = O.lookup(pyr, smooth, r)
    mask = np.ones_like(out, dtype=bool)
    mask[0, :, 1, 2] = False
    assert np.abs(out - ref)[mask].max() <= 1e-3 * scale
    assert np.all(out[0, :, 1, 2] == 0)  # NaN coordinate -> every tap out of range (same rule as the lookup)


@pytest.mark.parametrize("mode", MODES)
def test_five_levels(pkg, mode):
    """More levels than RAFT uses (the fused epilogue covers 4; deeper ones are pooled afterwards)."""
    B, D, H, W, L, r = 1, 32, 40, 48, 5, 2
    rng = np.random.default_rng(9)
    f1 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    coords = O.coords_grid(B, H, W) + rng.normal(0, 3.0, (B, 2, H, W)).astype(np.float32)
    ref = O.lookup(O.corr_pyramid(f1, f2, L, np.float64), coords, r)
    blk = pkg.CorrBlock(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()), L, r, build_mode=mode)
    assert rel_err(blk(torch.from_numpy(coords).to(dev())), ref) <= TOL[mode]


def test_cuda_graph_capture_replays_bit_identically(pkg):
    """Every entry point launches on the caller's stream without host synchronisation, so a whole refinement
    loop (build + lookups, and the on-the-fly variant) captures into a CUDA graph and replays the same bits."""
    d = dev()
    g = torch.Generator(device=d).manual_seed(21)
    B, D, H, W, L, r = 2, 128, 24, 40, 4, 4
    f1 = torch.randn((B, D, H, W), generator=g, device=d)
    f2 = torch.randn((B, D, H, W), generator=g, device=d)
    coords = [torch.from_numpy(O.coords_grid(B, H, W)).to(d) + 2.0 * torch.randn((B, 2, H, W), generator=g, device=d)
              for _ in range(3)]

    def loop():
        blk = pkg.CorrBlock(f1, f2, L, r)
        alt = pkg.AlternateCorrBlock(f1, f2, L, r)
        return [blk(c) for c in coords] + [alt(coords[0])]

    with torch.no_grad():
        eager = [o.clone() for o in loop()]
        side = torch.cuda.Stream(d)
        side.wait_stream(torch.cuda.current_stream(d))
        with torch.cuda.stream(side):
            loop()  # warm-up on the capture stream (lazy module loading is not capturable)
        torch.cuda.current_stream(d).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = loop()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
    for a, b in zip(eager, outs):
        assert torch.equal(a, b)


def test_two_devices_from_one_process(pkg):
    """nn.DataParallel (raft/train.py:71) drives several GPUs from threads of ONE process: every entry point takes the
    device explicitly, per-device state (function attributes, tensor maps) must not leak between devices."""
    import threading

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(3)
    B, D, H, W, L, r = 2, 64, 24, 40, 4, 4
    f1 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    coords = O.coords_grid(B, H, W) + rng.normal(0, 2.0, (B, 2, H, W)).astype(np.float32)
    results, errors = {}, []

    def work(i):
        try:
            d = torch.device(f"cuda:{i}")
            a = torch.from_numpy(f1).to(d).requires_grad_()
            b = torch.from_numpy(f2).to(d)
            c = torch.from_numpy(coords).to(d)
            outs = []
            for _ in range(3):
                blk = pkg.CorrBlock(a, b, L, r)
                o = blk(c)
                alt = pkg.AlternateCorrBlock(a.detach(), b, L, r)(c)
                outs.append((o.detach().cpu(), alt.cpu()))
            o.sum().backward()
            results[i] = (outs, a.grad.cpu())
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in (1, 0)]  # device 1 first: nothing may assume device 0
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    (o0, g0), (o1, g1) = results[0], results[1]
    for (a0, b0), (a1, b1) in zip(o0, o1):
        assert torch.equal(a0, a1) and torch.equal(b0, b1)
    assert torch.equal(o0[0][0], o0[2][0])
    ref = O.lookup(O.corr_pyramid(f1, f2, L, np.float64), coords, r)
    assert rel_err(o1[0][0], ref) <= TOL["tf32"]
    assert float((g0 - g1).abs().max()) <= 1e-5 * float(g0.abs().max())  # atomics: order differs between runs


@pytest.mark.parametrize("shape", [(2, 64, 55, 128, 4, 4), (3, 32, 23, 37, 4, 4), (1, 32, 46, 62, 4, 3), (2, 16, 17, 50, 3, 2)])
def test_lookup_tile_variants_agree_bit_for_bit(pkg, shape):
    """The quad-layout lookup runs with 32-pixel tiles (headline shape) or 16-pixel tiles (small / ragged problems), picked
    by shape; RAFTCORR_LOOKUP_TILE forces either.  Same boxes, same arithmetic per window: identical bits, and both
    within tolerance of the exact-fp32 build's lookup (corr.py:45-91)."""
    import os

    B, D, H, W, L, r = shape
    g = torch.Generator(device=dev()).manual_seed(17)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    ys, xs = torch.meshgrid(torch.arange(H, device=dev()), torch.arange(W, device=dev()), indexing="ij")
    coords = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1) + 5.0 * torch.randn((B, 2, H, W), generator=g, device=dev())
    coords[0, :, 0, 0] = torch.tensor([-20.0, 3.0], device=dev())   # a window entirely outside
    coords[-1, :, -1, -1] = torch.tensor([W + 1.5, H - 0.25], device=dev())
    blk = pkg.CorrBlock(f1, f2, L, r, build_mode="tf32")
    outs = {}
    old = os.environ.get("RAFTCORR_LOOKUP_TILE")
    try:
        for t in ("16", "32"):
            os.environ["RAFTCORR_LOOKUP_TILE"] = t
            outs[t] = blk(coords)
    finally:
        if old is None:
            os.environ.pop("RAFTCORR_LOOKUP_TILE", None)
        else:
            os.environ["RAFTCORR_LOOKUP_TILE"] = old
    assert torch.equal(outs["16"], outs["32"])
    assert torch.equal(blk(coords), outs["16"])  # whatever the launcher picks by itself
    ref = pkg.CorrBlock(f1, f2, L, r, build_mode="fp32")(coords)
    assert rel_err(outs["32"], ref.cpu().numpy()) <= TOL["tf32"]
