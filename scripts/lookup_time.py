"""Per-launch device time of the lookup (and the build) at the bench shape under a CUDA graph of 12 lookups.
usage: lookup_time.py [B [D H W r]]   env: RAFTCORR_PYRAMID_LAYOUT=quads|pairs|rowmajor"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
D, H, W, R = (int(a) for a in sys.argv[2:6]) if len(sys.argv) >= 6 else (256, 55, 128, 4)  # lookup_time.py B D H W r
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
coords = [(grid + flow).contiguous()]
for _ in range(11): coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
with torch.no_grad():
    blk = rc.CorrBlock(f1, f2, 4, R)
    ref = rc.CorrBlock(f1, f2, 4, R, build_mode="fp32")
    err = max(float((blk(c) - ref(c)).abs().max() / ref(c).abs().max()) for c in coords[:3])
    del ref
    for c in coords: blk(c)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for c in coords: o = blk(c)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record()
        for _ in range(5): gr.replay()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 60 * 1e3)
    tb = []
    for _ in range(5):
        e0.record()
        for _ in range(4): b2 = rc.CorrBlock(f1, f2, 4, R)
        e1.record(); torch.cuda.synchronize()
        tb.append(e0.elapsed_time(e1) / 4 * 1e3)
    # the whole inference step (build + 12 lookups) as ONE graph: what a batch-1 caller gets by capturing its loop
    gs = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gs):
        b3 = rc.CorrBlock(f1, f2, 4, R)
        for c in coords: o = b3(c)
    gs.replay(); torch.cuda.synchronize()
    tg = []
    for _ in range(5):
        e0.record()
        for _ in range(10): gs.replay()
        e1.record(); torch.cuda.synchronize()
        tg.append(e0.elapsed_time(e1) / 10 * 1e3)
print(f"shape B={B} D={D} {H}x{W} r={R}: build + 12 lookups as one CUDA graph {sorted(tg)[2]:.1f} us per step")
print(f"layout={os.environ.get('RAFTCORR_PYRAMID_LAYOUT','quads')} B={B}: lookup {sorted(ts)[2]:.1f} us (min {min(ts):.1f}) per launch (graph of 12), "
      f"build group {sorted(tb)[2]:.0f} us, max rel err vs fp32 mode {err:.2e}")
