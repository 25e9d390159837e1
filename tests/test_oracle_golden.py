"""Pin the oracles (numpy, plain C, torch-CPU port) against golden vectors produced by the
reference's own CorrBlock (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import corr_oracle as O
from make_golden import grad_weights

# fp32 re-orderings of the same arithmetic: the reference's grid_sample normalise/denormalise
# round trip (utils.py:74-78) alone moves coordinates by ~1e-5 px.
TOL_F32 = 2e-5
TOL_F64 = 2e-6  # fp64 oracle vs fp64 reference stored as fp32


def _meta(g):
    B, D, H, W, L, r = [int(v) for v in g["meta"]]
    return B, D, H, W, L, r


def test_numpy_pyramid_and_lookup_fp32(golden):
    B, D, H, W, L, r = _meta(golden)
    pyr = O.corr_pyramid(golden["fmap1"], golden["fmap2"], L, np.float32)
    assert [p.shape[-2:] for p in pyr] == O.level_shapes(H, W, L)
    for l in range(L):
        np.testing.assert_allclose(pyr[l], golden[f"pyr{l}"], atol=TOL_F32, rtol=0)
    for k in range(3):
        out = O.lookup(pyr, golden[f"coords{k}"], r)
        assert out.shape == (B, L * (2 * r + 1) ** 2, H, W)
        np.testing.assert_allclose(out, golden[f"out{k}"], atol=TOL_F32, rtol=0)


def test_numpy_fp64_matches_reference_fp64(golden):
    B, D, H, W, L, r = _meta(golden)
    pyr = O.corr_pyramid(golden["fmap1"], golden["fmap2"], L, np.float64)
    for k in range(3):
        out = O.lookup(pyr, golden[f"coords{k}"], r)
        np.testing.assert_allclose(out, golden[f"out64_{k}"], atol=TOL_F64, rtol=0)


def test_numpy_gradients(golden):
    B, D, H, W, L, r = _meta(golden)
    pyr = O.corr_pyramid(golden["fmap1"], golden["fmap2"], L, np.float64)
    acc = None
    for k in range(3):
        go = grad_weights(golden[f"out{k}"].shape, k).astype(np.float32).astype(np.float64)
        d, dco = O.lookup_grad(pyr, golden[f"coords{k}"], go, r, want_coords=True)
        acc = d if acc is None else [a + b for a, b in zip(acc, d)]
        if k == 1:  # generic (non-integer) coordinates: the derivative is unique there
            np.testing.assert_allclose(dco, golden["dcoords1"], atol=2e-5, rtol=0)
    df1, df2 = O.build_grad(acc, golden["fmap1"], golden["fmap2"])
    np.testing.assert_allclose(df1, golden["df1"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(df2, golden["df2"], atol=5e-6, rtol=0)


def test_numpy_alt_equals_reference_corrblock(golden):
    """AlternateCorrBlock == CorrBlock by linearity of pooling; the on-the-fly
    restatement is pinned by the same golden outputs and gradients."""
    B, D, H, W, L, r = _meta(golden)
    a1 = a2 = 0
    for k in range(3):
        out = O.alt_lookup(golden["fmap1"], golden["fmap2"], golden[f"coords{k}"], L, r, np.float64)
        np.testing.assert_allclose(out, golden[f"out64_{k}"], atol=TOL_F64, rtol=0)
        go = grad_weights(golden[f"out{k}"].shape, k).astype(np.float32).astype(np.float64)
        x, y = O.alt_grad(golden["fmap1"], golden["fmap2"], golden[f"coords{k}"], go, L, r, np.float64)
        a1, a2 = a1 + x, a2 + y
    np.testing.assert_allclose(a1, golden["df1"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(a2, golden["df2"], atol=5e-6, rtol=0)


def test_c_oracle(golden):
    from oracle import c_oracle as C

    B, D, H, W, L, r = _meta(golden)
    pyr = C.corr_pyramid(golden["fmap1"], golden["fmap2"], L)
    for l in range(L):
        np.testing.assert_allclose(pyr[l], golden[f"pyr{l}"], atol=TOL_F32, rtol=0)
    for k in range(3):
        np.testing.assert_allclose(C.lookup(pyr, golden[f"coords{k}"], r), golden[f"out{k}"], atol=TOL_F32, rtol=0)
        np.testing.assert_allclose(C.alt_lookup(golden["fmap1"], golden["fmap2"], golden[f"coords{k}"], L, r),
                                   golden[f"out{k}"], atol=TOL_F32, rtol=0)
    go = grad_weights(golden["out1"].shape, 1).astype(np.float32)
    d = C.lookup_grad([p.shape for p in pyr], golden["coords1"], go, r)
    dn, _ = O.lookup_grad([p.astype(np.float64) for p in pyr], golden["coords1"], go.astype(np.float64), r)
    for a, b in zip(d, dn):
        np.testing.assert_allclose(a, b, atol=TOL_F32, rtol=0)


def test_torch_cpu_port(golden):
    import torch

    from oracle.torch_cpu_port import CpuCorrBlock

    B, D, H, W, L, r = _meta(golden)
    blk = CpuCorrBlock(torch.from_numpy(golden["fmap1"]), torch.from_numpy(golden["fmap2"]), L, r)
    for l in range(L):
        np.testing.assert_allclose(blk.corr_pyramid[l].numpy().reshape(golden[f"pyr{l}"].shape), golden[f"pyr{l}"],
                                   atol=TOL_F32, rtol=0)
    for k in range(3):
        out = blk(torch.from_numpy(golden[f"coords{k}"])).numpy()
        np.testing.assert_allclose(out, golden[f"out{k}"], atol=TOL_F32, rtol=0)


def test_pool_adjoint_is_adjoint():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 7, 9))
    y = rng.normal(size=(3, 3, 4))
    assert abs((O.avg_pool2(x) * y).sum() - (x * O.avg_pool2_adjoint(y, 7, 9)).sum()) < 1e-12


def test_numpy_lookup_convc1_matches_reference(golden_convc1):
    """SURVEY 8(f) rank 1 oracle: lookup -> convc1 -> ReLU against the reference CorrBlock + the reference motion
    encoder's own convc1 (raft/core/update.py:135,170), fp64 and fp32."""
    g = golden_convc1
    B, D, H, W, L, r, cout = [int(v) for v in g["meta"]]
    for dt, tol in ((np.float64, TOL_F64), (np.float32, 1e-5)):
        pyr = O.corr_pyramid(g["fmap1"], g["fmap2"], L, dt)
        for k in range(3):
            out = O.lookup_convc1(pyr, g[f"coords{k}"], g["weight"], g["bias"], r)
            assert out.shape == (B, cout, H, W)
            ref = g[f"cor64_{k}"].astype(np.float64)
            assert np.abs(out - ref).max() <= tol * np.abs(ref).max(), (dt, k)
            assert float(g[f"fp32_err{k}"]) <= 1e-5  # the reference's own fp32 run agrees with its fp64 run


def test_numpy_fnet_head_matches_reference(golden_fnet_head):
    """SURVEY 8(f) rank 2 oracle: conv2 head + split against the reference encoders' own ``conv2`` on the input a
    forward hook captured inside ``self.fnet([image_1, image_2])`` (raft/core/extractor.py:231,239 / :341,349), then
    the volume and lookups of the reference CorrBlock on those maps."""
    g = golden_fnet_head
    B, Cin, D, H, W, L, r = [int(v) for v in g["meta"]]
    assert float(g["fp32_err_head"]) <= 1e-5
    for dt, tol in ((np.float64, 1e-6), (np.float32, 1e-5)):  # goldens are fp64 answers stored as fp32
        f1, f2 = O.fnet_head(g["x"], g["weight"], g["bias"], dt)
        assert f1.shape == (B, D, H, W) and f2.shape == (B, D, H, W)
        for f, ref in ((f1, g["fmap1"]), (f2, g["fmap2"])):
            assert np.abs(f - ref.astype(np.float64)).max() <= tol * np.abs(ref).max(), dt
        pyr = O.corr_pyramid(f1, f2, L, dt)
        assert np.abs(pyr[0] - g["pyr0"].astype(np.float64)).max() <= tol * np.abs(g["pyr0"]).max()
        for k in range(3):
            out = O.lookup(pyr, g[f"coords{k}"], r)
            ref = g[f"out64_{k}"].astype(np.float64)
            assert np.abs(out - ref).max() <= tol * np.abs(ref).max(), (dt, k)
            assert float(g[f"fp32_err{k}"]) <= 1e-5


def test_numpy_upsample_flow_matches_reference(golden_upsample):
    """SURVEY 8(f) rank 4 oracle: convex upsampling and its adjoint against the reference's RAFT.upsample_flow
    (raft/core/raft.py:70-81) and torch autograd, fp64."""
    g = golden_upsample
    flow, mask = g["flow"].astype(np.float64), g["mask"].astype(np.float64)
    up = O.upsample_flow(flow, mask)
    assert up.shape == g["up64"].shape
    assert np.abs(up - g["up64"]).max() <= TOL_F64 * np.abs(g["up64"]).max()
    from make_golden import grad_weights
    dflow, dmask = O.upsample_flow_grad(flow, mask, grad_weights(up.shape, 0))
    assert np.abs(dflow - g["dflow64"]).max() <= TOL_F64 * np.abs(g["dflow64"]).max()
    assert np.abs(dmask - g["dmask64"]).max() <= TOL_F64 * np.abs(g["dmask64"]).max()
    up32 = O.upsample_flow(g["flow"], g["mask"])
    assert up32.dtype == np.float32 and np.abs(up32 - g["up64"]).max() <= 1e-5 * np.abs(g["up64"]).max()
