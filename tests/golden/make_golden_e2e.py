"""End-to-end golden: the reference's CorrBlock + SmallUpdateBlock refinement loop on seeded inputs.

Run in the build container only.  Checks that tests/raft_harness.py reproduces the reference's
update operator bit-for-bit-ish, records the reference flows of all 12 iterations, and measures how
much of the 0.01 px EPE budget plain arithmetic noise already uses (fp64 run, TF32-rounded inputs).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.environ.get("RAFT_REFERENCE", "/root/reference/raft"))

import raft_harness as RH  # noqa: E402
from core.corr import CorrBlock  # noqa: E402  (the reference, unmodified)
from core.update import SmallUpdateBlock  # noqa: E402


def tf32_round(t):
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF  # round-to-nearest (ties away) on the 13 dropped mantissa bits
    return i.view(torch.float32)


def epe(a, b):
    return float((a - b).pow(2).sum(1).sqrt().mean())


def main():
    torch.set_num_threads(4)
    op = RH.seeded_operator()
    ref_op = SmallUpdateBlock(argparse.Namespace(corr_levels=RH.LEVELS, corr_radius=RH.RADIUS), hidden_dim=RH.HIDDEN).eval()
    sd = {}
    for mine, theirs in RH.REFERENCE_KEYS.items():
        # the reference ConvGRU names its convs convz / convr / convq exactly (update.py:60-62)
        sd[theirs + ".weight"] = getattr(op, mine).weight.detach().clone()
        sd[theirs + ".bias"] = getattr(op, mine).bias.detach().clone()
    ref_op.load_state_dict(sd, strict=True)
    f1, f2, net, inp = RH.seeded_inputs()

    class RefOp(torch.nn.Module):
        def forward(self, net, inp, corr, flow):
            n, _, d = ref_op(net, inp, corr, flow)
            return n, d

    with torch.no_grad():
        ref = RH.refine(CorrBlock, RefOp(), f1, f2, net, inp)
        mine = RH.refine(CorrBlock, op, f1, f2, net, inp)
        d_harness = max(epe(a, b) for a, b in zip(ref, mine))
        ref64 = RH.refine(CorrBlock, op.double(), f1.double(), f2.double(), net.double(), inp.double())
        op.float()
        tf = RH.refine(CorrBlock, op, tf32_round(f1), tf32_round(f2), net, inp)
    print(f"harness vs reference update operator: max EPE over iterations {d_harness:.2e}")
    print(f"fp32 vs fp64 reference loop: final EPE {epe(ref[-1], ref64[-1].float()):.2e}")
    print(f"TF32-rounded feature maps (what the tensor-core build computes): final EPE {epe(ref[-1], tf[-1]):.2e}")
    print(f"|flow| mean {float(ref[-1].abs().mean()):.3f} max {float(ref[-1].abs().max()):.3f}")
    assert d_harness < 1e-5
    np.savez_compressed(os.path.join(HERE, "e2e_raft_small.npz"), flows=torch.stack(ref).numpy(),
                        flows64=torch.stack(ref64).float().numpy())
    print("wrote e2e_raft_small.npz", os.path.getsize(os.path.join(HERE, "e2e_raft_small.npz")))


if __name__ == "__main__":
    main()
