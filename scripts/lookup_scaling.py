"""Where does the TMA lookup's time go?  Vary radius (rows per box) and level count (boxes per pixel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W = 8, 256, 55, 128
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
coords = [(grid + 4 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous() for _ in range(8)]
with torch.no_grad():
    for L in (4, 2, 1):
        for r in (4, 3, 2, 1):
            blk = rc.CorrBlock(f1, f2, L, r)
            for c in coords: blk(c)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                for c in coords: o = blk(c)
            e1.record(); torch.cuda.synchronize()
            K = 2 * r + 1
            us = e0.elapsed_time(e1) / 32 * 1e3
            print(f"L={L} r={r}: {us:6.1f} us/lookup  boxes/px={L} rows/box={K+1} out MB={B*L*K*K*H*W*4/1e6:.1f}")
