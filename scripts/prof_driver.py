"""Small driver for ncu captures: one f16 build + 3 lookups at the bench shape (B=8, 55x128, D=256, r=4, L=4), then
(with `alt`) 2 on-the-fly lookups at 1080p B=4, then (with `bwd`) a config-3 backward.  usage: prof_driver.py [alt] [bwd]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc

dev = torch.device("cuda:0")


def case(B, D, H, W, n, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev)
    f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
    flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
    coords = [(grid + flow).contiguous()]
    for _ in range(n - 1):
        coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
    return f1, f2, coords


with torch.no_grad():
    f1, f2, coords = case(8, 256, 55, 128, 3)
    for _ in range(2):
        blk = rc.CorrBlock(f1, f2, 4, 4)
        for c in coords:
            o = blk(c)
    torch.cuda.synchronize()
    del blk, o
    if "alt" in sys.argv:
        f1, f2, coords = case(4, 256, 135, 240, 2, seed=1)
        ab = rc.AlternateCorrBlock(f1, f2, 4, 4)
        for _ in range(2):
            for c in coords:
                o = ab(c)
        torch.cuda.synchronize()
        del ab, o
if "bwd" in sys.argv:
    f1, f2, coords = case(10, 256, 46, 62, 3, seed=2)
    f1.requires_grad_(); f2.requires_grad_()
    for _ in range(2):
        blk = rc.CorrBlock(f1, f2, 4, 4)
        loss = 0
        for c in coords:
            loss = loss + blk(c).square().sum()
        loss.backward()
        f1.grad = f2.grad = None
    torch.cuda.synchronize()
print("prof_driver done")
