"""Generate golden vectors for the RAFT correlation path by running the UNMODIFIED reference.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

It imports ``core.corr.CorrBlock`` from ``/root/reference/raft`` (raft/core/corr.py:13-116),
runs it on seeded inputs in fp32 and fp64 on the CPU, and records inputs + outputs:

* ``pyr{l}``         the reference ``corr_pyramid`` levels, fp32           (corr.py:31-43)
* ``out{k}``         ``corr_fn(coords_k)``, fp32 and (``out64_{k}``) fp64   (corr.py:45-91)
* ``df1/df2``        autograd gradients of ``sum_k <out_k, w_k>`` w.r.t. the feature maps, fp64
* ``dcoords{k}``     autograd gradient w.r.t. the coordinates, fp64
  (``w_k`` is the deterministic weight field :func:`grad_weights`, recomputed by the tests)

Nothing from the reference is copied: only its numerical outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = os.environ.get("RAFT_REFERENCE", "/root/reference/raft")
HERE = os.path.dirname(os.path.abspath(__file__))


def grad_weights(shape, k):
    """Deterministic upstream-gradient field shared by generator and tests."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.float64)
    return (np.sin(0.37 * idx + 1.3 * k) + 0.25 * np.cos(0.011 * idx)).reshape(shape)


def make_coords(rng, B, H, W, kind):
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    grid = np.stack([xs, ys], 0).astype(np.float32)[None].repeat(B, 0)
    if kind == "grid":  # exact integers: the first RAFT iteration (raft.py:130)
        return grid
    if kind == "noise":
        return grid + rng.normal(0, 4.0, grid.shape).astype(np.float32)
    if kind == "adversarial":  # half-integers, negatives, far out of range, level edges
        c = grid.copy()
        flat = c.reshape(B, 2, -1)
        n = flat.shape[-1]
        sel = rng.permutation(n)
        q = n // 8
        flat[:, :, sel[0 * q:1 * q]] += 0.5
        flat[:, :, sel[1 * q:2 * q]] -= 0.5
        flat[:, 0, sel[2 * q:3 * q]] = -3.25
        flat[:, 1, sel[3 * q:4 * q]] = -0.75
        flat[:, 0, sel[4 * q:5 * q]] = W - 1 + 0.5
        flat[:, 1, sel[5 * q:6 * q]] = H + 7.0
        flat[:, :, sel[6 * q:7 * q]] = -100.0
        flat[:, 0, sel[7 * q:]] = W - 1.0
        return c
    raise ValueError(kind)


CASES = {
    # name: (B, D, H, W, num_levels, radius, seed)
    "small_r3_odd": (1, 32, 19, 22, 4, 3, 11),   # RAFT-small-like: r=3, odd sizes -> floor pooling
    "basic_r4": (1, 24, 16, 20, 4, 4, 12),       # RAFT-basic-like: r=4
    "levels3_r4": (2, 16, 9, 12, 3, 4, 13),      # num_levels != default
}


def main():
    import torch

    sys.path.insert(0, REF)
    from core.corr import CorrBlock  # the reference, unmodified

    torch.set_num_threads(1)
    for name, (B, D, H, W, L, r, seed) in CASES.items():
        rng = np.random.default_rng(seed)
        f1 = rng.normal(0, 1, (B, D, H, W)).astype(np.float32)
        f2 = rng.normal(0, 1, (B, D, H, W)).astype(np.float32)
        coords = [make_coords(rng, B, H, W, k) for k in ("grid", "noise", "adversarial")]
        rec = {"fmap1": f1, "fmap2": f2, "meta": np.array([B, D, H, W, L, r], np.int64)}
        for k, c in enumerate(coords):
            rec[f"coords{k}"] = c

        # fp32 forward, exactly as raft.py:119,144 would call it
        with torch.no_grad():
            blk = CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), num_levels=L, radius=r)
            for l, p in enumerate(blk.corr_pyramid):
                rec[f"pyr{l}"] = p.numpy().reshape(B, H * W, p.shape[-2], p.shape[-1]).copy()
            for k, c in enumerate(coords):
                rec[f"out{k}"] = blk(torch.from_numpy(c)).numpy().copy()

        # fp64 forward + autograd (the "true" values; CorrBlock.__call__ ends in .float(), corr.py:91,
        # so gradients are taken through an fp64 re-run with that cast being a no-op on the graph)
        t1 = torch.from_numpy(f1).double().requires_grad_()
        t2 = torch.from_numpy(f2).double().requires_grad_()
        blk64 = CorrBlock(t1, t2, num_levels=L, radius=r)
        loss = 0
        cts = []
        for k, c in enumerate(coords):
            ct = torch.from_numpy(c).double().requires_grad_()
            cts.append(ct)
            # re-implement only the final cast: call the reference up to corr.py:90 via __call__ and
            # accept its .float() (values) -- gradients flow through the cast back to fp64 leaves.
            o = blk64(ct)
            rec[f"out64_{k}"] = o.detach().numpy().astype(np.float32)
            wk = torch.from_numpy(grad_weights(tuple(o.shape), k)).to(o.dtype)
            loss = loss + (o * wk).sum()
        loss.backward()
        rec["df1"] = t1.grad.numpy().astype(np.float32)
        rec["df2"] = t2.grad.numpy().astype(np.float32)
        for k, ct in enumerate(cts):
            rec[f"dcoords{k}"] = ct.grad.numpy().astype(np.float32)
        path = os.path.join(HERE, f"corr_{name}.npz")
        np.savez_compressed(path, **rec)
        print(name, {k: v.shape for k, v in rec.items() if k.startswith(("pyr", "out0"))},
              f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
