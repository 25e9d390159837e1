"""Is the TMA lookup bound by DRAM bytes?  Time it with window starts that make every level-0 box row fall in
ONE 64-byte block (x0 % 16 == 0), in TWO (x0 % 16 == 12), and at random; same for all levels as far as possible."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W, r = 8, 256, 55, 128, 4
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
blk = rc.CorrBlock(f1, f2, 4, r)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)

def timed(make):
    cs = [make(i).contiguous() for i in range(8)]
    with torch.no_grad():
        for c in cs: blk(c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            for c in cs: blk(c)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 48 * 1e3

def snapped(phase, scale):
    # x such that floor(x / scale) - r == phase (mod 16) at the level with that scale, random otherwise
    def make(i):
        c = grid + 6 * torch.randn((B, 2, H, W), generator=g, device=dev)
        x = c[:, 0]
        k = torch.floor((x / scale - r - phase) / 16)
        xl = (k * 16 + phase + r + 0.37) * scale
        return torch.stack([xl, c[:, 1]], 1)
    return make

print(f"random            : {timed(lambda i: grid + 6 * torch.randn((B, 2, H, W), generator=g, device=dev)):.1f} us")
for scale in (1, 2, 4):
    for phase in (0, 4, 8, 12):
        print(f"level {scale:d}x start%16={phase:2d}: {timed(snapped(phase, scale)):.1f} us")
