"""On-the-fly lookup at small sizes: tensor-core blocks vs exact-fp32 CUDA-core kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
for (B, D, H, W, L, r) in [(1, 128, 46, 62, 4, 3), (1, 256, 46, 62, 4, 4), (1, 256, 55, 128, 4, 4), (2, 256, 55, 128, 4, 4),
                           (8, 256, 55, 128, 4, 4), (1, 256, 135, 240, 4, 4)]:
    g = torch.Generator(device=dev).manual_seed(1)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
    flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
    coords = (grid + flow).contiguous()
    res = []
    for mode in ("tf32", "fp32"):
        with torch.no_grad():
            blk = rc.AlternateCorrBlock(f1, f2, L, r, build_mode=mode)
            for _ in range(3): blk(coords)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()   # take the host out of the measurement
            with torch.cuda.graph(graph):
                for _ in range(10): o = blk(coords)
            graph.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 10 * 1e3)
    print(f"B={B} D={D} {H}x{W} L={L} r={r}: tensor-core {res[0]:8.1f} us   cuda-core fp32 {res[1]:8.1f} us")
