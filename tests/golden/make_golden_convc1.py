"""Golden vectors for the fused lookup + convc1 + ReLU, produced by the UNMODIFIED reference.

Run in the build container only:

    python tests/golden/make_golden_convc1.py      # writes tests/golden/convc1_*.npz

For RAFT-small- and RAFT-basic-shaped cases it instantiates the reference's own motion encoders
(``SmallMotionEncoder`` / ``BasicMotionEncoder``, raft/core/update.py:105-177) with seeded random weights, runs the
reference ``CorrBlock`` (raft/core/corr.py:13-116) and records what the first line of the encoder's ``forward``
computes from it -- ``cor = F.relu(self.convc1(corr))`` (update.py:135 / :170) -- in fp64 on the CPU (plus how far the
reference's own fp32 run is from that), together with the inputs and ``convc1``'s weight and bias.  Only numerical outputs are stored.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = os.environ.get("RAFT_REFERENCE", "/root/reference/raft")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import make_coords  # noqa: E402

CASES = {
    # name: (small, B, D, H, W, seed)      corr_levels = 4; radius 3 (small) / 4 (basic), raft.py:30-38
    "small": (True, 1, 32, 19, 22, 21),    # convc1: 196 -> 96
    "basic": (False, 1, 24, 16, 20, 22),   # convc1: 324 -> 256
}


def main():
    import torch
    import torch.nn.functional as F

    sys.path.insert(0, REF)
    from core.corr import CorrBlock
    from core.update import BasicMotionEncoder, SmallMotionEncoder

    torch.set_num_threads(1)
    for name, (small, B, D, H, W, seed) in CASES.items():
        args = argparse.Namespace(corr_levels=4, corr_radius=3 if small else 4)
        torch.manual_seed(seed)
        enc = (SmallMotionEncoder if small else BasicMotionEncoder)(args)
        rng = np.random.default_rng(seed)
        f1 = rng.normal(0, 1, (B, D, H, W)).astype(np.float32)
        f2 = rng.normal(0, 1, (B, D, H, W)).astype(np.float32)
        rec = {"meta": np.array([B, D, H, W, args.corr_levels, args.corr_radius, enc.convc1.out_channels], np.int64),
               "fmap1": f1, "fmap2": f2,
               "weight": enc.convc1.weight.detach().numpy().copy(), "bias": enc.convc1.bias.detach().numpy().copy()}
        for k, kind in enumerate(["grid", "noise", "adversarial"]):
            coords = make_coords(rng, B, H, W, kind)
            rec[f"coords{k}"] = coords
            with torch.no_grad():
                corr = CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), radius=args.corr_radius)(torch.from_numpy(coords))
                cor32 = F.relu(enc.convc1(corr))                           # update.py:135 / :170, as the reference runs it
                enc64 = enc.double()
                corr64 = CorrBlock(torch.from_numpy(f1).double(), torch.from_numpy(f2).double(),
                                   radius=args.corr_radius)(torch.from_numpy(coords).double()).double()
                cor64 = F.relu(enc64.convc1(corr64))
                enc.float()
                # the fp64 evaluation is the recorded answer (stored as fp32: 6e-8 relative, far below the tolerance);
                # the reference's own fp32 run differs from it by this much
                rec[f"cor64_{k}"] = cor64.float().numpy()
                rec[f"fp32_err{k}"] = np.array(float((cor32.double() - cor64).abs().max() / cor64.abs().max()))
        path = os.path.join(HERE, f"convc1_{name}.npz")
        np.savez_compressed(path, **rec)
        print(path, {k: v.shape for k, v in rec.items() if k.startswith("cor")})


if __name__ == "__main__":
    main()
