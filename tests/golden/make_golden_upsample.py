"""Golden vectors for the convex 8x flow upsampling, produced by the UNMODIFIED reference.

Run in the build container only:

    python tests/golden/make_golden_upsample.py     # writes tests/golden/upsample_*.npz

It calls the reference's own ``RAFT.upsample_flow`` (raft/core/raft.py:70-81; the method does not touch ``self``) in
fp64 on seeded inputs and records the result and its autograd gradients w.r.t. flow and mask for the deterministic
upstream gradient ``grad_weights`` (tests/golden/make_golden.py).  Only numerical outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = os.environ.get("RAFT_REFERENCE", "/root/reference/raft")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import grad_weights  # noqa: E402

CASES = {
    # name: (B, C, H, W, seed)
    "flow": (1, 2, 5, 37, 31),          # ragged: 37 = 32 + 5 pixels per row
    "flow_var": (1, 4, 6, 16, 32),      # flow + flow variance in one pass (uncertainty model, raft.py:165-168)
}


def main():
    import torch

    sys.path.insert(0, REF)
    from core.raft import RAFT

    torch.set_num_threads(1)
    for name, (B, C, H, W, seed) in CASES.items():
        rng = np.random.default_rng(seed)
        flow = rng.normal(0, 3.0, (B, C, H, W)).astype(np.float32)
        mask = rng.normal(0, 2.0, (B, 576, H, W)).astype(np.float32)
        f = torch.from_numpy(flow).double().requires_grad_()
        m = torch.from_numpy(mask).double().requires_grad_()
        # the reference handles C = 2; more channels go through it two at a time, exactly like raft.py:165-168
        ups = [RAFT.upsample_flow(None, f[:, c:c + 2], m) for c in range(0, C, 2)]
        up = torch.cat(ups, dim=1)
        gw = torch.from_numpy(grad_weights(tuple(up.shape), 0))
        (up * gw).sum().backward()
        rec = {"meta": np.array([B, C, H, W], np.int64), "flow": flow, "mask": mask,
               "up64": up.detach().float().numpy(), "dflow64": f.grad.float().numpy(), "dmask64": m.grad.float().numpy()}
        path = os.path.join(HERE, f"upsample_{name}.npz")
        np.savez_compressed(path, **rec)
        print(path, {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
