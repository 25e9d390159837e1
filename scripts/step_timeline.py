"""Event timeline of bench steps: where do the microseconds between kernels go? (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W, L, r, iters = 8, 256, 55, 128, 4, 4, 12
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
coords = [(grid + 4 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous() for _ in range(iters)]
for mode in sys.argv[1:] or ["tma", "cpasync"]:
    os.environ["RAFTCORR_LOOKUP"] = mode
    for _ in range(3):
        blk = rc.CorrBlock(f1, f2, L, r)
        for c in coords: o = blk(c)
    torch.cuda.synchronize()
    steps = 4
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(iters + 2)] for _ in range(steps)]
    for s in range(steps):
        ev[s][0].record()
        blk = rc.CorrBlock(f1, f2, L, r)
        ev[s][1].record()
        for i, c in enumerate(coords):
            o = blk(c)
            ev[s][2 + i].record()
    torch.cuda.synchronize()
    for s in range(steps):
        t = [ev[s][i].elapsed_time(ev[s][i + 1]) * 1e3 for i in range(iters + 1)]
        nxt = ev[s][-1].elapsed_time(ev[s + 1][0]) * 1e3 if s + 1 < steps else 0
        print(f"{mode} step {s}: build {t[0]:.0f} us | lookups " + " ".join(f"{x:.0f}" for x in t[1:]) + f" | gap to next step {nxt:.0f} | total {sum(t):.0f}")
