"""A few fused lookup+convc1 launches at the bench shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W, L, r, cout = 8, 256, 55, 128, 4, 4, 256
f1 = torch.randn((B, D, H, W), device=dev); f2 = torch.randn((B, D, H, W), device=dev)
w = torch.randn((cout, 324, 1, 1), device=dev) / 18; bias = torch.randn((cout,), device=dev)
yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
base = torch.stack([xx, yy]).float()[None].repeat(B, 1, 1, 1)
flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), device=dev), 5, 1, 2)
with torch.no_grad():
    blk = rc.CorrBlock(f1, f2, L, r)
    for i in range(4):
        out = blk.lookup_convc1(base + flow + 0.5 * i * torch.randn((B, 2, H, W), device=dev), w, bias)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
