"""Golden vectors for the encoder output head fused into the build, produced by the
UNMODIFIED reference.

Run in the build container only:

    python tests/golden/make_golden_fnet_head.py      # writes tests/golden/fnet_head_*.npz

It instantiates the reference's own feature encoders (``BasicEncoder(output_dim=256, norm_fn='instance')`` /
``SmallEncoder(output_dim=128, norm_fn='instance')``, raft/core/raft.py:41-44,48-51) with seeded random weights, runs
them in eval mode on a seeded frame pair exactly as RAFT does (``self.fnet([image_1, image_2])``, raft.py:110-111) and
records, through a forward hook on ``conv2`` (raft/core/extractor.py:172/287, called at :231/:341):

* ``x``                the head's input (output of ``layer3``), ``[2B, Cin, H, W]`` fp32
* ``weight, bias``     ``conv2``'s parameters
* ``fmap1, fmap2``     ``conv2(x)`` split per frame (:239/:349), evaluated in fp64 (stored fp32)
* ``pyr0``             level 0 of the reference ``CorrBlock``'s pyramid on those maps, fp64 arithmetic (corr.py:31-35)
* ``out64_{k}``        ``corr_fn(coords_k)`` in fp64 for the three coordinate sets of make_golden.py
* ``fp32_err_*``       how far the reference's own fp32 run is from the fp64 answer (sizes the tolerances)

Only numerical outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = os.environ.get("RAFT_REFERENCE", "/root/reference/raft")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import make_coords  # noqa: E402

CASES = {
    # name: (small, B, image H, image W, radius, seed)
    "basic": (False, 1, 128, 160, 4, 31),   # 16x20 maps, conv2 128 -> 256
    "small": (True, 2, 128, 144, 3, 32),    # 16x18 maps (288 px: a ragged third tile), conv2 96 -> 128, two pairs
}


def main():
    import torch

    sys.path.insert(0, REF)
    from core.corr import CorrBlock
    from core.extractor import BasicEncoder, SmallEncoder

    torch.set_num_threads(4)
    for name, (small, B, IH, IW, radius, seed) in CASES.items():
        torch.manual_seed(seed)
        enc = (SmallEncoder(output_dim=128, norm_fn="instance", dropout=0.0) if small
               else BasicEncoder(output_dim=256, norm_fn="instance", dropout=0.0)).eval()
        rng = np.random.default_rng(seed)
        # smooth random frames in [-1, 1] (raft.py:100-101 normalises to that range)
        img = rng.normal(0, 1, (2, B, 3, IH // 8, IW // 8)).astype(np.float32)
        img = torch.nn.functional.interpolate(torch.from_numpy(img).reshape(2 * B, 3, IH // 8, IW // 8), size=(IH, IW),
                                              mode="bicubic", align_corners=False).clamp(-1, 1)
        img = img + 0.1 * torch.from_numpy(rng.normal(0, 1, img.shape).astype(np.float32))
        image_1, image_2 = img[:B].contiguous(), img[B:].contiguous()
        grabbed = {}
        h = enc.conv2.register_forward_hook(lambda m, inp, out: grabbed.__setitem__("x", inp[0].detach().clone()))
        with torch.no_grad():
            fmap_1, fmap_2 = enc([image_1, image_2])   # raft.py:110-111
        h.remove()
        x = grabbed["x"]
        w, b = enc.conv2.weight.detach(), enc.conv2.bias.detach()
        with torch.no_grad():
            f64 = torch.nn.functional.conv2d(x.double(), w.double(), b.double())
            f1_64, f2_64 = torch.split(f64, [B, B], dim=0)
            head_fp32_err = float((torch.cat([fmap_1, fmap_2]).double() - f64).abs().max() / f64.abs().max())
            blk64 = CorrBlock(f1_64, f2_64, radius=radius)
            blk32 = CorrBlock(fmap_1.float(), fmap_2.float(), radius=radius)   # raft.py:114-119
            H, W = f64.shape[-2:]
            rec = {"meta": np.array([B, x.shape[1], w.shape[0], H, W, 4, radius], np.int64),
                   "x": x.numpy(), "weight": w.numpy().copy(), "bias": b.numpy().copy(),
                   "fmap1": f1_64.float().numpy(), "fmap2": f2_64.float().numpy(),
                   "pyr0": blk64.corr_pyramid[0].reshape(B, H * W, H, W).float().numpy(),
                   "fp32_err_head": np.array(head_fp32_err)}
            for k, kind in enumerate(["grid", "noise", "adversarial"]):
                coords = make_coords(rng, B, H, W, kind)
                rec[f"coords{k}"] = coords
                o64 = blk64(torch.from_numpy(coords).double()).double()
                o32 = blk32(torch.from_numpy(coords))
                rec[f"out64_{k}"] = o64.float().numpy()
                rec[f"fp32_err{k}"] = np.array(float((o32.double() - o64).abs().max() / o64.abs().max()))
        path = os.path.join(HERE, f"fnet_head_{name}.npz")
        np.savez_compressed(path, **rec)
        print(path, {k: (v.shape if v.ndim else float(v)) for k, v in rec.items()})


if __name__ == "__main__":
    main()
