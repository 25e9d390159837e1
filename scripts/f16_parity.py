"""Volume / lookup error of the three build modes against fp64 at the bench shape (one pair), same inputs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W, L, r = 1, 256, 55, 128, 4, 4
N = H * W
g = torch.Generator(device=dev).manual_seed(1234)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
coords = (torch.stack([xs, ys]).float()[None] + 4 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous()
ref = torch.einsum("dm,dn->mn", f1[0].reshape(D, N).double(), f2[0].reshape(D, N).double()) / D ** 0.5
out = {}
with torch.no_grad():
    lk_ref = rc.CorrBlock(f1, f2, L, r, build_mode="fp32")(coords).double()
    for mode in ("fp32", "tf32", "f16"):
        blk = rc.CorrBlock(f1, f2, L, r, build_mode=mode)
        v0 = blk.corr_pyramid[0].reshape(N, N).double()
        e = (v0 - ref).abs()
        lk = blk(coords).double()
        out[mode] = {"volume_max_err_over_max": float(e.max() / ref.abs().max()), "volume_rms_err_over_rms": float((e.pow(2).mean().sqrt()) / ref.pow(2).mean().sqrt()),
                     "lookup_max_err_over_max_vs_fp32_build": float((lk - lk_ref).abs().max() / lk_ref.abs().max())}
print(json.dumps(out))
