import os, sys
sys.path.insert(0, "/root/repo")
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W = 8, 256, 55, 128
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
coords = [(grid + flow).contiguous()]
for _ in range(11): coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
with torch.no_grad():
    blk = rc.CorrBlock(f1, f2, 4, 4)
    for c in coords: blk(c)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for c in coords: o = blk(c)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
print(f"EXP={os.environ.get('RAFTCORR_LOOKUP_EXP','0')}: {e0.elapsed_time(e1)/60*1e3:.1f} us per lookup (CUDA graph of 12, bench coords)")
