"""Encoder head fused into the build: `CorrBlock.from_encoder_head(x, conv2.weight, conv2.bias)` vs
the reference sequence conv2 (cuDNN) -> split -> CorrBlock(fmap1, fmap2), config-2 / 3 / 4 shapes.  CUDA events,
several input sets rotated so that consecutive calls do not find x in L2.  One JSON line per shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")

def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for (B, Cin, D, H, W, r) in [(8, 128, 256, 55, 128, 4), (10, 128, 256, 46, 62, 4), (16, 128, 256, 47, 156, 4), (1, 96, 128, 46, 62, 3)]:
    xs = [torch.randn((2 * B, Cin, H, W), device=dev) for _ in range(3)]
    w = torch.randn((D, Cin, 1, 1), device=dev) / Cin ** 0.5
    b = 0.1 * torch.randn((D,), device=dev)
    k = [0]
    def fused():
        k[0] += 1
        return rc.CorrBlock.from_encoder_head(xs[k[0] % 3], w, b, radius=r)
    def unfused():
        k[0] += 1
        f = F.conv2d(xs[k[0] % 3], w, b)
        return rc.CorrBlock(f[:B], f[B:], radius=r)
    def conv_only():
        k[0] += 1
        return F.conv2d(xs[k[0] % 3], w, b)
    fm = F.conv2d(xs[0], w, b)
    def build_only():
        return rc.CorrBlock(fm[:B], fm[B:], radius=r)
    with torch.no_grad():
        res = {"B": B, "Cin": Cin, "D": D, "H": H, "W": W,
               "fused_head_build_us": t(fused), "conv2_cudnn_plus_build_us": t(unfused), "conv2_cudnn_us": t(conv_only),
               "build_us": t(build_only)}
        coords = torch.zeros((B, 2, H, W), device=dev)
        k[0] = 0
        a = fused()(coords)
        k[0] = 0
        c = unfused()(coords)
        res["lookup_rel_diff_fused_vs_unfused"] = float((a - c).abs().max() / c.abs().max())
    res["head_share_us"] = res["fused_head_build_us"] - res["build_us"]
    res["head_algorithmic_MB"] = (2 * B * Cin * H * W * 4 + 2 * B * D * H * W * 4) / 1e6
    print(json.dumps(res), flush=True)
