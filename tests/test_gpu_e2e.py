"""End-to-end drop-in check (north_star: <= 0.01 px final-flow EPE delta after 12 iterations).

The RAFT-small refinement loop (tests/raft_harness.py; proven identical to the reference's
SmallUpdateBlock in tests/golden/make_golden_e2e.py) runs on the GPU with THIS repo's CorrBlock /
AlternateCorrBlock and is compared with the flows the reference produced with its own CorrBlock."""
import os

import numpy as np
import pytest
import torch

import raft_harness as RH

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "e2e_raft_small.npz")
EPE_BUDGET = 0.01  # px at 1/8 resolution, mean end-point error of the final flow


def _epe(a, b):
    return float(np.sqrt(((a - b) ** 2).sum(1)).mean())


@pytest.mark.parametrize("variant", ["corr_f16", "corr_f16x3", "corr_tf32", "corr_fp32", "alternate", "corr_f16_fusedconvc1", "corr_tf32_fusedconvc1", "corr_fp32_fusedconvc1"])
def test_flow_after_12_iterations_matches_reference(variant):
    import optical_flow_dnn_b200 as rc

    torch.backends.cudnn.allow_tf32 = False  # isolate the correlation path: the update operator runs in fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    gold = np.load(GOLD)["flows"]
    op = RH.seeded_operator().to(dev)
    f1, f2, net, inp = (t.to(dev) for t in RH.seeded_inputs())
    with torch.no_grad():
        if variant == "alternate":
            flows = RH.refine(rc.AlternateCorrBlock, op, f1, f2, net, inp)
        else:
            flows = RH.refine(rc.CorrBlock, op, f1, f2, net, inp, build_mode=variant.split("_")[1],
                              fused_convc1=variant.endswith("fusedconvc1"))
    got = torch.stack(flows).cpu().numpy()
    assert got.shape == gold.shape
    per_iter = [_epe(got[i], gold[i]) for i in range(len(gold))]
    print(variant, "EPE per iteration:", " ".join(f"{e:.1e}" for e in per_iter))
    assert per_iter[-1] <= EPE_BUDGET, per_iter
    assert max(per_iter) <= EPE_BUDGET, per_iter


def test_training_step_gradients_flow_through_12_iterations():
    """Config-3-style use: backward through every lookup and the volume (one shared gradient pyramid)."""
    import optical_flow_dnn_b200 as rc
    from oracle import corr_oracle as O

    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    B, D, H, W, L, r, iters = 2, 32, 16, 20, 4, 3, 4
    f1 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    coords = [O.coords_grid(B, H, W) + rng.normal(0, 2.0, (B, 2, H, W)).astype(np.float32) for _ in range(iters)]
    gws = [rng.normal(size=(B, L * (2 * r + 1) ** 2, H, W)).astype(np.float32) for _ in range(iters)]
    # oracle: sum over iterations of the lookup adjoint, then the build adjoint (fp64)
    pyr = O.corr_pyramid(f1, f2, L, np.float64)
    acc = None
    for c, gw in zip(coords, gws):
        d, _ = O.lookup_grad(pyr, c, gw.astype(np.float64), r)
        acc = d if acc is None else [a + b for a, b in zip(acc, d)]
    e1, e2 = O.build_grad(acc, f1, f2)
    for cls, kw in ((rc.CorrBlock, {"build_mode": "fp32"}), (rc.CorrBlock, {"build_mode": "tf32"}), (rc.CorrBlock, {"build_mode": "f16"}),
                    (rc.AlternateCorrBlock, {})):
        t1 = torch.from_numpy(f1).to(dev).requires_grad_()
        t2 = torch.from_numpy(f2).to(dev).requires_grad_()
        blk = cls(t1, t2, num_levels=L, radius=r, **kw)
        loss = 0
        for c, gw in zip(coords, gws):
            loss = loss + (blk(torch.from_numpy(c).to(dev)) * torch.from_numpy(gw).to(dev)).sum()
        loss.backward()
        for got, ref in ((t1.grad, e1), (t2.grad, e2)):
            err = float(np.abs(got.cpu().numpy() - ref).max() / np.abs(ref).max())
            assert err <= 1e-3, (cls.__name__, kw, err)
        # a second, independent graph on the same inputs must give the same gradients (fresh accumulator)
        g1 = t1.grad.clone()
        t1.grad = t2.grad = None
        blk = cls(t1, t2, num_levels=L, radius=r, **kw)
        loss = 0
        for c, gw in zip(coords, gws):
            loss = loss + (blk(torch.from_numpy(c).to(dev)) * torch.from_numpy(gw).to(dev)).sum()
        loss.backward()
        assert float((t1.grad - g1).abs().max()) <= 1e-4 * float(g1.abs().max())
