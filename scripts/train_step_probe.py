import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W, L, r, iters = 10, 256, 46, 62, 4, 4, 12
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
coords = [(grid + 2 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous() for _ in range(iters)]
ws = [torch.randn((B, L * 81, H, W), device=dev) for _ in range(iters)]
for rep in range(3):
    a = f1.clone().requires_grad_(); b = f2.clone().requires_grad_()
    blk = rc.CorrBlock(a, b, L, r)
    torch.autograd.backward([blk(c) for c in coords], ws)
torch.cuda.synchronize()
print("ok")
