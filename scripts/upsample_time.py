"""Convex upsampling: this library vs the reference recipe in stock PyTorch, config-2 and config-4 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")

def recipe(flow, mask):
    N, C, H, W = flow.shape
    mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, C, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, C, 8 * H, 8 * W)

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for (B, C, H, W) in [(8, 2, 55, 128), (16, 2, 47, 156), (16, 4, 47, 156), (1, 2, 55, 128)]:
    flow = 4 * torch.randn((B, C, H, W), device=dev); mask = 2 * torch.randn((B, 576, H, W), device=dev)
    # several masks so that consecutive calls do not hit L2 (B=8: 130 MB each)
    masks = [mask.clone() for _ in range(4)]
    k = [0]
    def ours():
        k[0] += 1
        return rc.upsample_flow(flow, masks[k[0] & 3])
    def stock():
        k[0] += 1
        return recipe(flow, masks[k[0] & 3])
    with torch.no_grad():
        to, ts = t(ours), t(stock)
    byts = B * H * W * (576 * 4 + C * 4 + C * 64 * 4)
    print(f"B={B} C={C} {H}x{W}: ours {to:.1f} us ({byts / to / 1e3:.0f} GB/s algorithmic) | stock PyTorch recipe {ts:.1f} us | {ts / to:.1f}x")
    fg = flow.clone().requires_grad_(); mg = masks[0].clone().requires_grad_()
    gout = torch.randn((B, C, 8 * H, 8 * W), device=dev)
    def ours_fb():
        fg.grad = mg.grad = None
        rc.upsample_flow(fg, mg).backward(gout)
    def stock_fb():
        fg.grad = mg.grad = None
        recipe(fg, mg).backward(gout)
    print(f"    forward + backward: ours {t(ours_fb, 10):.1f} us | stock {t(stock_fb, 10):.1f} us")
