"""The reference's UNMODIFIED caller (`raft/core/raft.py:83-187`, `update.py`, `extractor.py`, `utils/utils.py`; staged
byte-identical under baseline/_ref by scripts/vendor_reference.py) run twice on the GPU with the same seeded weights
and frames: once with the reference's own `core/corr.py` (stock PyTorch CorrBlock, fp32, TF32 off) and once with
`core/corr.py` replaced by the INTEGRATION.md 2a re-export of this library.  north_star: "raft.py, the RAFT-small and
raft_uncertainty models, train.py and evaluate.py pick it up unchanged", final-flow EPE delta <= 0.01 px after 12
iterations.  Frames are padded and the model called exactly as `raft/evaluate.py:114-120` does.
"""
import numpy as np
import pytest
import torch

import ref_loader as RL

pytestmark = pytest.mark.gpu
# Parameter gradients of a whole training step (4 refinement iterations, GRU in the loop), max|delta| / max|reference
# gradient| per tensor.  The operator-level gradient tolerance is 1e-3 (tests/test_gpu_grad_train_shape.py, measured
# 3.4e-4); through the recurrent update block the tensor-core modes' 3e-4 volume rounding also moves the iterates
# themselves, measured 2.4e-3 .. 5.6e-3 on the encoder weights.  The exact-fp32 build mode shows the floor of such a
# comparison: cuBLAS SGEMM vs this library's FFMA build differ by summation order only (1e-6), and that alone moves the
# same gradients by 3e-4 .. 7e-4.
GRAD_TOL = {"f16": 1e-2, "fp32": 3e-3}
EPE_BUDGET = 0.01  # px at full resolution, mean end-point error of the final (upsampled) flow, 12 iterations

MODELS = {
    "small": dict(small=True),                                  # RAFT-small: D=128, r=3          (raft.py:29-34,47-50)
    "basic": dict(small=False),                                 # RAFT basic: D=256, r=4          (raft.py:35-39,51-54)
    "uncertainty": dict(small=False, uncertainty=True),         # raft_uncertainty head           (update.py:224-227)
    "uncertainty_residual": dict(small=False, uncertainty=True, residual_variance=True, log_variance=True),
    "alternate": dict(small=False, alternate_corr=True),        # --alternate_corr                (raft.py:116-117)
    "small_alternate": dict(small=True, alternate_corr=True),
}
SHAPES = {"chairs_368x496": (368, 496), "sintel_436x1024": (436, 1024), "kitti_375x1242": (375, 1242)}


def _frames(h, w, dev, seed, batch=1):
    """A textured frame and a warped + re-noised copy: real correspondences, sub-pixel motion of a few pixels."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(batch, 3, h // 4 + 8, w // 4 + 8, generator=g)
    big = torch.nn.functional.interpolate(base, scale_factor=4, mode="bicubic", align_corners=False)
    big = (big + 0.25 * torch.rand(batch, 3, big.shape[2], big.shape[3], generator=g)).clamp(0, 1.25) / 1.25
    im1 = big[:, :, 12:12 + h, 12:12 + w]
    im2 = big[:, :, 9:9 + h, 17:17 + w] + 0.02 * torch.randn(batch, 3, h, w, generator=g)
    return (255 * im1).clamp(0, 255).to(dev).contiguous(), (255 * im2).clamp(0, 255).to(dev).contiguous()


def _models(kind, dev, seed=11):
    if not RL.available():
        pytest.fail("baseline/_ref is not staged: run `python -c 'import __graft_entry__ as g; g.build()'` in the build "
                    "container (needs /root/reference)")
    stock, dropin = RL.load_arm("stock"), RL.load_arm("dropin")
    import optical_flow_dnn_b200 as rc

    assert dropin.CorrBlock is rc.CorrBlock and dropin.AlternateCorrBlock is rc.AlternateCorrBlock
    assert stock.CorrBlock is not rc.CorrBlock and stock.CorrBlock.__module__ == "core.corr"
    torch.manual_seed(seed)
    ref = stock.RAFT(RL.raft_args(**MODELS[kind])).to(dev)
    new = dropin.RAFT(RL.raft_args(**MODELS[kind])).to(dev)
    new.load_state_dict(ref.state_dict())
    return stock, ref, new


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False  # the stock arm is the reference's fp32 arithmetic
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False


def _epe(a, b):
    return float((a - b).pow(2).sum(1).sqrt().mean())


CASES = [("small", "chairs_368x496"), ("small", "sintel_436x1024"), ("basic", "chairs_368x496"), ("basic", "sintel_436x1024"),
         ("uncertainty", "sintel_436x1024"), ("uncertainty", "kitti_375x1242"), ("uncertainty_residual", "chairs_368x496"),
         ("alternate", "chairs_368x496"), ("alternate", "sintel_436x1024"), ("small_alternate", "chairs_368x496")]


@pytest.mark.parametrize("kind,shape", CASES)
def test_unmodified_raft_forward_flow_epe(kind, shape):
    _no_tf32()
    dev = torch.device("cuda:0")
    stock, ref, new = _models(kind, dev)
    ref.eval(), new.eval()
    h, w = SHAPES[shape]
    im1, im2 = _frames(h, w, dev, seed=5)
    # raft/evaluate.py:114-120: pad to a multiple of 8, 12 refinement iterations here (north_star), test_mode
    pad = stock._utils_module.InputPadder(im1.shape)  # raft/core/utils/utils.py:7-27
    p1, p2 = pad.pad(im1, im2)
    with torch.no_grad():
        if MODELS[kind].get("alternate_corr"):
            # the primary reference is the stock CorrBlock (AlternateCorrBlock == CorrBlock) ...
            ref.args.alternate_corr = False
            a_ref, f_ref = ref(p1, p2, iters=12, test_mode=True)
            ref.args.alternate_corr = True
        else:
            a_ref, f_ref = ref(p1, p2, iters=12, test_mode=True)
        a_new, f_new = new(p1, p2, iters=12, test_mode=True)
    f_ref, f_new = pad.unpad(f_ref), pad.unpad(f_new)
    assert f_new.shape == (1, 2, h, w) and torch.isfinite(f_new).all()
    epe = _epe(f_new, f_ref)
    mag = float(f_ref.pow(2).sum(1).sqrt().mean())
    print(f"{kind} {shape}: flow EPE(drop-in vs reference CorrBlock) = {epe:.2e} px  (mean |flow| {mag:.2f} px)")
    assert epe <= EPE_BUDGET, epe
    if MODELS[kind].get("uncertainty"):
        v_ref, v_new = pad.unpad(a_ref), pad.unpad(a_new)
        dv = float((v_new - v_ref).abs().max() / v_ref.abs().max().clamp_min(1e-12))
        print(f"{kind} {shape}: flow-variance max rel-to-max delta = {dv:.2e}")
        assert dv <= 1e-2, dv
    if MODELS[kind].get("alternate_corr") and RL.reference_alt_extension_available():
        # ... and the reference's own AlternateCorrBlock over its own compiled alt_cuda_corr kernel as second oracle
        with torch.no_grad():
            _, f_alt = ref(p1, p2, iters=12, test_mode=True)
        epe_alt = _epe(f_new, pad.unpad(f_alt))
        print(f"{kind} {shape}: flow EPE(drop-in vs reference AlternateCorrBlock/alt_cuda_corr) = {epe_alt:.2e} px")
        assert epe_alt <= EPE_BUDGET, epe_alt


@pytest.mark.parametrize("mode", ["fp32", "tf32", "f16x3"])
def test_unmodified_raft_forward_other_build_modes(mode):
    """The same run with the library's other build modes (selected the way a user would: set_default_build_mode)."""
    import optical_flow_dnn_b200 as rc
    from optical_flow_dnn_b200 import corr as rc_corr

    _no_tf32()
    dev = torch.device("cuda:0")
    stock, ref, new = _models("basic", dev)
    ref.eval(), new.eval()
    im1, im2 = _frames(368, 496, dev, seed=6)
    before = rc_corr._default_mode
    rc.set_default_build_mode(mode)
    try:
        with torch.no_grad():
            _, f_ref = ref(im1, im2, iters=12, test_mode=True)
            _, f_new = new(im1, im2, iters=12, test_mode=True)
    finally:
        rc.set_default_build_mode(before)
    epe = _epe(f_new, f_ref)
    print(f"basic 368x496 build_mode={mode}: flow EPE = {epe:.2e} px")
    assert epe <= EPE_BUDGET, epe


@pytest.mark.parametrize("mode", ["f16", "fp32"])
@pytest.mark.parametrize("kind", ["small", "basic", "uncertainty"])
def test_unmodified_raft_training_step_parameter_gradients(kind, mode):
    """train.py's step (`raft/train.py:118-128`: forward in train mode, a loss over all iterations' predictions,
    backward) through the unmodified RAFT.forward: every parameter gradient of the drop-in arm -- the feature encoder's
    come ONLY through this library's lookup and build backward -- against the stock arm's."""
    import optical_flow_dnn_b200 as rc
    from optical_flow_dnn_b200 import corr as rc_corr

    _no_tf32()
    dev = torch.device("cuda:0")
    stock, ref, new = _models(kind, dev)
    ref.train(), new.train()
    before = rc_corr._default_mode
    rc.set_default_build_mode(mode)
    for m in (ref, new):
        m.freeze_bn()  # train.py:84-85 (stage != chairs)
    im1, im2 = _frames(368, 496, dev, seed=8, batch=2)
    g = torch.Generator().manual_seed(3)
    target = (4 * torch.randn(2, 2, 368, 496, generator=g)).to(dev)
    losses = []
    try:
        for model in (ref, new):
            preds = model(im1, im2, iters=4)
            loss = 0
            for i, p in enumerate(preds):  # the reference's sequence loss shape (losses.py: gamma-weighted L1 over iterations)
                flow = p[0] if isinstance(p, tuple) else p
                loss = loss + 0.8 ** (len(preds) - 1 - i) * (flow - target).abs().mean()
                if isinstance(p, tuple):
                    loss = loss + 0.1 * 0.8 ** (len(preds) - 1 - i) * p[1].abs().mean()
            loss.backward()
            losses.append(float(loss))
    finally:
        rc.set_default_build_mode(before)
    assert abs(losses[0] - losses[1]) <= 1e-4 * abs(losses[0]), losses
    # Biases in front of an instance / batch norm (fnet.conv1.bias ...) have a mathematically ZERO gradient: what the two
    # arms hold there is rounding noise (1e-9 of the other gradients), so every tensor is measured against
    # max(its own max|grad|, 1e-2 x the largest gradient magnitude
    # ... of its own top-level module: fnet / cnet / update_block).
    gmax = {}
    for name, p in ref.named_parameters():
        if p.grad is not None:
            top = name.split(".")[0]
            gmax[top] = max(gmax.get(top, 0.0), float(p.grad.abs().max()))
    rows = []
    for (name, p_ref), (_, p_new) in zip(ref.named_parameters(), new.named_parameters()):
        if p_ref.grad is None:
            assert p_new.grad is None, name
            continue
        assert p_new.grad is not None, name
        own = float(p_ref.grad.abs().max())
        scale = max(own, 1e-2 * gmax[name.split(".")[0]])
        rows.append((float((p_new.grad - p_ref.grad).abs().max()) / scale, name, own))
    rows.sort(reverse=True)
    n_fnet = sum(1 for _, name, own in rows if name.startswith("fnet.") and own > 1e-2 * gmax["fnet"])
    print(f"{kind} build_mode={mode}: per-module max|grad| {gmax}; {n_fnet} fnet tensors compared on their own scale; worst five:",
          [(f"{e:.1e}", n) for e, n, _ in rows[:5]])
    worst = (rows[0][1], rows[0][0])
    assert n_fnet >= 10, n_fnet  # the feature encoder's gradients (the ones that cross this library) are really compared
    assert worst[1] <= GRAD_TOL[mode], rows[:5]
    print(f"{kind}: training-step loss {losses[0]:.6f} vs {losses[1]:.6f}; worst parameter-gradient delta {worst[1]:.2e} ({worst[0]})")
