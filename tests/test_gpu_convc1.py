"""SURVEY.md 8(f) rank 1 on the GPU: the lookup fused with the motion encoder's first 1x1 convolution + ReLU
(`raftcorr_lookup_convc1_fwd`, called through the C ABI) against

  (1) golden vectors from the reference CorrBlock + the reference motion encoders' own ``convc1``
      (tests/golden/make_golden_convc1.py; raft/core/update.py:135,170),
  (2) the CPU oracle on seeded shapes (ragged tiles, fewer levels, every radius, Cout = 16 .. 256),
  (3) at BASELINE.json's full size: the unfused pair ``convc1(lookup(coords))`` evaluated in fp64 on the device.

Tolerance: max|delta| <= 2e-3 * max|ref| per tensor.  Two TF32 stages stack here (north_star's 1e-3 budget for the
volume/lookup, plus the 324-term TF32 product with fp32 accumulation -- the precision cuDNN uses for this convolution
by default); measured 2e-4 .. 5e-4.  With an fp32-built volume only the product is rounded: 1e-3.
"""
import numpy as np
import pytest
import torch

from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu
TOL = {"tf32": 2e-3, "fp32": 1e-3}


def dev():
    return torch.device("cuda:0")


def rel_err(a, ref):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    return float(np.abs(a.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.fixture(scope="module")
def pkg():
    import optical_flow_dnn_b200 as m

    return m


@pytest.mark.parametrize("mode", ["tf32", "fp32"])
def test_golden_reference_motion_encoder(pkg, golden_convc1, mode):
    g = golden_convc1
    B, D, H, W, L, r, cout = [int(v) for v in g["meta"]]
    w = torch.from_numpy(g["weight"]).to(dev())
    bias = torch.from_numpy(g["bias"]).to(dev())
    with torch.no_grad():
        blk = pkg.CorrBlock(torch.from_numpy(g["fmap1"]).to(dev()), torch.from_numpy(g["fmap2"]).to(dev()), L, r, build_mode=mode)
        for k in range(3):
            out = blk.lookup_convc1(torch.from_numpy(g[f"coords{k}"]).to(dev()), w, bias)
            assert tuple(out.shape) == (B, cout, H, W) and out.dtype == torch.float32 and out.is_contiguous()
            assert rel_err(out, g[f"cor64_{k}"].astype(np.float64)) <= TOL[mode], f"coords set {k}"


SEEDED = [
    # B, D, H, W, L, r, Cout
    (1, 64, 20, 28, 4, 4, 256),   # 560 px: ragged last tile
    (2, 48, 19, 37, 3, 2, 48),    # odd sizes, 3 levels, one N half
    (1, 32, 8, 16, 1, 1, 16),     # one level, smallest everything: exactly one tile
    (3, 32, 16, 24, 4, 3, 96),    # RAFT-small channel counts, three tiles per image
    (1, 32, 17, 23, 2, 4, 192),   # two N halves of 96
]


@pytest.mark.parametrize("shape", SEEDED)
@pytest.mark.parametrize("mode", ["tf32", "fp32"])
def test_seeded_vs_oracle(pkg, shape, mode):
    B, D, H, W, L, r, cout = shape
    rng = np.random.default_rng(77)
    f1 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    cin = L * (2 * r + 1) ** 2
    w = (rng.normal(size=(cout, cin, 1, 1)) / np.sqrt(cin)).astype(np.float32)
    bias = rng.normal(size=(cout,)).astype(np.float32)
    pyr = O.corr_pyramid(f1, f2, L, np.float64)
    with torch.no_grad():
        blk = pkg.CorrBlock(torch.from_numpy(f1).to(dev()), torch.from_numpy(f2).to(dev()), L, r, build_mode=mode)
        coords = O.coords_grid(B, H, W)
        for it in range(3):
            coords = coords + rng.normal(0, 4.0 if it == 0 else 0.5, coords.shape).astype(np.float32)
            c = torch.from_numpy(coords).to(dev())
            ref = O.lookup_convc1(pyr, coords, w, bias, r)
            assert rel_err(blk.lookup_convc1(c, torch.from_numpy(w).to(dev()), torch.from_numpy(bias).to(dev())), ref) <= TOL[mode]
            ref_lin = O.lookup_convc1(pyr, coords, w, None, r, relu=False)
            assert rel_err(blk.lookup_convc1(c, torch.from_numpy(w).to(dev()), None, relu=False), ref_lin) <= TOL[mode]


def test_full_size_against_unfused(pkg):
    """436x1024 -> 55x128, D=256, r=4, Cout=256 (RAFT-basic, BASELINE.json configs[1] shape), B=2."""
    B, D, H, W, L, r, cout = 2, 256, 55, 128, 4, 4, 256
    g = torch.Generator(device=dev()).manual_seed(99)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    w = torch.randn((cout, L * 81, 1, 1), generator=g, device=dev()) / 18.0
    bias = torch.randn((cout,), generator=g, device=dev())
    coords = torch.from_numpy(O.coords_grid(B, H, W)).to(dev()) + 4.0 * torch.randn((B, 2, H, W), generator=g, device=dev())
    with torch.no_grad():
        blk = pkg.CorrBlock(f1, f2, L, r)
        feat = blk(coords)
        ref = torch.relu(torch.nn.functional.conv2d(feat.double(), w.double(), bias.double()))
        out = blk.lookup_convc1(coords, w, bias)
        assert float((out.double() - ref).abs().max() / ref.abs().max()) <= 1e-3  # same volume: only the product differs
        # deterministic, and the packed weight follows in-place parameter updates
        assert torch.equal(out, blk.lookup_convc1(coords, w, bias))
        w.mul_(2.0)
        out2 = blk.lookup_convc1(coords, w, None, relu=False)
        ref2 = torch.nn.functional.conv2d(feat.double(), w.double())
        assert float((out2.double() - ref2).abs().max() / ref2.abs().max()) <= 1e-3


@pytest.mark.parametrize("relu", [True, False])
def test_autograd_matches_unfused_path(pkg, relu):
    """Training use: gradients w.r.t. the feature maps, convc1's weight and bias through the fused call (backward
    recomputes the lookup, nothing of size [B, 324, H, W] is kept) against torch autograd through lookup + conv2d."""
    B, D, H, W, L, r, cout = 2, 64, 20, 28, 4, 3, 96
    g = torch.Generator(device=dev()).manual_seed(17)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev())
    f2 = torch.randn((B, D, H, W), generator=g, device=dev())
    cin = L * (2 * r + 1) ** 2
    w = torch.randn((cout, cin, 1, 1), generator=g, device=dev()) / cin ** 0.5
    b = torch.randn((cout,), generator=g, device=dev())
    coords = [torch.from_numpy(O.coords_grid(B, H, W)).to(dev()) + s * torch.randn((B, 2, H, W), generator=g, device=dev())
              for s in (0.0, 3.0)]
    gout = [torch.randn((B, cout, H, W), generator=g, device=dev()) for _ in coords]

    def run(fused):
        a1, a2, ww, bb = (z.detach().clone().requires_grad_(True) for z in (f1, f2, w, b))
        blk = pkg.CorrBlock(a1, a2, L, r, build_mode="fp32")
        loss = 0
        for c, go in zip(coords, gout):
            if fused:
                o = blk.lookup_convc1(c, ww, bb, relu=relu)
            else:
                o = torch.nn.functional.conv2d(blk(c).double(), ww.double(), bb.double())
                if relu:  # same ReLU mask in both arms: an activation within TF32 rounding of zero may flip otherwise,
                    with torch.no_grad():  # and one flipped unit moves a whole pixel's gradient by O(1)
                        mask = blk.lookup_convc1(c, ww.detach(), bb.detach(), relu=True) > 0
                    o = o * mask
                o = o.float()
            loss = loss + (o * go).sum()
        loss.backward()
        return a1.grad, a2.grad, ww.grad, bb.grad

    got, want = run(True), run(False)
    for name, a, ref in zip(("dfmap1", "dfmap2", "dweight", "dbias"), got, want):
        assert a is not None and a.shape == ref.shape, name
        assert float((a - ref).abs().max() / ref.abs().max()) <= 2e-3, name


def test_errors(pkg):
    d = dev()
    f = torch.randn(1, 32, 16, 16, device=d)
    blk = pkg.CorrBlock(f, f, 2, 3)
    c = torch.zeros(1, 2, 16, 16, device=d)
    with pytest.raises(RuntimeError):
        blk.lookup_convc1(c, torch.randn(96, 99, device=d))          # wrong Cin
    with pytest.raises(pkg.RaftCorrError):
        with torch.no_grad():
            blk.lookup_convc1(c, torch.randn(40, 98, device=d))      # Cout not a multiple of 16
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            blk.lookup_convc1(c, torch.randn(96, 98))                # CPU weight
