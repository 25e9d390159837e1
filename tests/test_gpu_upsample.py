"""SURVEY.md 8(f) rank 4 on the GPU: convex 8x flow upsampling (`raftcorr_upsample_flow_fwd/_bwd` through the C ABI)
against (1) the reference's own RAFT.upsample_flow + torch autograd (goldens), (2) the CPU oracle on seeded shapes,
(3) the reference recipe in stock PyTorch fp64 on the device at BASELINE.json's sizes.
fp32 arithmetic throughout: tolerance 1e-5 of max|ref| (forward), 1e-4 (gradients: atomics re-order the flow sums)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from make_golden import grad_weights
from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def rel_err(a, ref):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    return float(np.abs(a.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.fixture(scope="module")
def pkg():
    import optical_flow_dnn_b200 as m

    return m


def test_golden_reference_upsample_flow(pkg, golden_upsample):
    g = golden_upsample
    flow = torch.from_numpy(g["flow"]).to(dev()).requires_grad_()
    mask = torch.from_numpy(g["mask"]).to(dev()).requires_grad_()
    up = pkg.upsample_flow(flow, mask)
    assert tuple(up.shape) == g["up64"].shape and up.dtype == torch.float32 and up.is_contiguous()
    assert rel_err(up, g["up64"].astype(np.float64)) <= 1e-5
    gw = torch.from_numpy(grad_weights(tuple(up.shape), 0).astype(np.float32)).to(dev())
    (up * gw).sum().backward()
    assert rel_err(flow.grad, g["dflow64"].astype(np.float64)) <= 1e-4
    assert rel_err(mask.grad, g["dmask64"].astype(np.float64)) <= 1e-4


@pytest.mark.parametrize("shape", [(1, 2, 1, 1), (2, 2, 7, 33), (1, 1, 16, 64), (3, 3, 5, 31), (1, 4, 23, 78), (2, 2, 46, 62)])
def test_seeded_vs_oracle(pkg, shape):
    B, C, H, W = shape
    rng = np.random.default_rng(5)
    flow = rng.normal(0, 3.0, (B, C, H, W)).astype(np.float32)
    mask = rng.normal(0, 3.0, (B, 576, H, W)).astype(np.float32)
    gout = rng.normal(size=(B, C, 8 * H, 8 * W)).astype(np.float32)
    ref = O.upsample_flow(flow.astype(np.float64), mask.astype(np.float64))
    dflow, dmask = O.upsample_flow_grad(flow.astype(np.float64), mask.astype(np.float64), gout.astype(np.float64))
    f = torch.from_numpy(flow).to(dev()).requires_grad_()
    m = torch.from_numpy(mask).to(dev()).requires_grad_()
    up = pkg.upsample_flow(f, m)
    assert rel_err(up, ref) <= 1e-5
    up.backward(torch.from_numpy(gout).to(dev()))
    assert rel_err(f.grad, dflow) <= 1e-4
    assert rel_err(m.grad, dmask) <= 1e-4
    with torch.no_grad():  # inference path (no autograd node) gives the same bits
        assert torch.equal(pkg.upsample_flow(f.detach(), m.detach()), up.detach())


def _reference_recipe(flow, mask):
    """raft/core/raft.py:70-81 in stock PyTorch (fp64 on the device)."""
    N, C, H, W = flow.shape
    mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, C, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, C, 8 * H, 8 * W)


@pytest.mark.parametrize("shape", [(8, 2, 55, 128), (4, 4, 47, 156)])
def test_full_size_against_stock_pytorch(pkg, shape):
    B, C, H, W = shape
    g = torch.Generator(device=dev()).manual_seed(17)
    flow = (4.0 * torch.randn((B, C, H, W), generator=g, device=dev())).requires_grad_()
    mask = (2.0 * torch.randn((B, 576, H, W), generator=g, device=dev())).requires_grad_()
    gout = torch.randn((B, C, 8 * H, 8 * W), generator=g, device=dev())
    up = pkg.upsample_flow(flow, mask)
    up.backward(gout)
    f64 = flow.detach().double().requires_grad_()
    m64 = mask.detach().double().requires_grad_()
    ref = _reference_recipe(f64, m64)
    ref.backward(gout.double())
    assert float((up.double() - ref).abs().max() / ref.abs().max()) <= 1e-5
    assert float((flow.grad.double() - f64.grad).abs().max() / f64.grad.abs().max()) <= 1e-4
    assert float((mask.grad.double() - m64.grad).abs().max() / m64.grad.abs().max()) <= 1e-4


def test_errors(pkg):
    d = dev()
    with pytest.raises(RuntimeError):
        pkg.upsample_flow(torch.zeros(1, 2, 4, 4), torch.zeros(1, 576, 4, 4))                      # CPU tensors
    with pytest.raises(RuntimeError):
        pkg.upsample_flow(torch.zeros(1, 2, 4, 4, device=d), torch.zeros(1, 64, 4, 4, device=d))   # wrong mask channels
    with pytest.raises(RuntimeError):
        pkg.upsample_flow(torch.zeros(1, 5, 4, 4, device=d), torch.zeros(1, 576, 4, 4, device=d))  # too many channels
    out = pkg.upsample_flow(torch.zeros(0, 2, 4, 4, device=d), torch.zeros(0, 576, 4, 4, device=d))
    assert tuple(out.shape) == (0, 2, 32, 32)
