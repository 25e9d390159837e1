"""Fused lookup + convc1 + ReLU against lookup -> torch conv (fp64) on the device; timing vs the unfused pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
torch.backends.cudnn.allow_tf32 = True
shapes = [(1, 64, 20, 28, 4, 4, 256), (2, 128, 46, 62, 4, 3, 96), (2, 64, 19, 37, 3, 2, 48), (1, 32, 8, 16, 1, 1, 16),
          (2, 256, 55, 128, 4, 4, 256)]
if len(sys.argv) > 1: shapes = shapes[: int(sys.argv[1])]
for (B, D, H, W, L, r, cout) in shapes:
    f1 = torch.randn((B, D, H, W), generator=g, device=dev)
    f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    cin = L * (2 * r + 1) ** 2
    w = torch.randn((cout, cin, 1, 1), generator=g, device=dev) / cin ** 0.5
    bias = torch.randn((cout,), generator=g, device=dev)
    yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    coords = torch.stack([xx, yy]).float()[None].repeat(B, 1, 1, 1) + 3.0 * torch.randn((B, 2, H, W), generator=g, device=dev)
    for mode in ("tf32", "fp32"):
        with torch.no_grad():
            blk = rc.CorrBlock(f1, f2, L, r, build_mode=mode)
            feat = blk(coords)
            ref = torch.relu(torch.nn.functional.conv2d(feat.double(), w.double(), bias.double()))
            out = blk.lookup_convc1(coords, w, bias)
            torch.cuda.synchronize()
            err = float((out.double() - ref).abs().max() / ref.abs().max())
            out2 = blk.lookup_convc1(coords, w, None, relu=False)
            ref2 = torch.nn.functional.conv2d(feat.double(), w.double())
            err2 = float((out2.double() - ref2).abs().max() / ref2.abs().max())
        print(f"B{B} D{D} {H}x{W} L{L} r{r} Cout{cout} {mode}: fused vs fp64 conv of our lookup: {err:.3e}  (no bias/relu {err2:.3e})", flush=True)

# timing at the bench shape
B, D, H, W, L, r, cout = 8, 256, 55, 128, 4, 4, 256
f1 = torch.randn((B, D, H, W), device=dev); f2 = torch.randn((B, D, H, W), device=dev)
w = torch.randn((cout, 324, 1, 1), device=dev) / 18; bias = torch.randn((cout,), device=dev)
yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
base = torch.stack([xx, yy]).float()[None].repeat(B, 1, 1, 1)
flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), device=dev), 5, 1, 2)
cs = [base + flow + 0.5 * i * torch.randn((B, 2, H, W), device=dev) for i in range(12)]
with torch.no_grad():
    blk = rc.CorrBlock(f1, f2, L, r)
    def t(fn, n=5):
        for c in cs: fn(c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for c in cs: fn(c)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / len(cs) * 1e3)
        return best
    t_lookup = t(lambda c: blk(c))
    t_unfused = t(lambda c: torch.relu_(torch.nn.functional.conv2d(blk(c), w, bias)))
    t_fused = t(lambda c: blk.lookup_convc1(c, w, bias))
print(f"B=8 55x128 r=4 Cout=256 per iteration: lookup alone {t_lookup:.1f} us | lookup + cuDNN conv1x1 + relu {t_unfused:.1f} us | fused {t_fused:.1f} us")
