"""Build group (operand prep + build_tc_kernel) at several shapes: us and GB/s of pyramid written.
usage: build_shape_time.py B,H,W [B,H,W ...]   (D = 256, 4 levels, r = 4, default build mode)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
D, LV = int(os.environ.get("FEAT", "256")), 4
for spec in sys.argv[1:] or ["8,55,128", "16,47,156"]:
    B, H, W = (int(x) for x in spec.split(","))
    f1 = torch.randn((B, D, H, W), device=dev); f2 = torch.randn((B, D, H, W), device=dev)
    with torch.no_grad():
        for _ in range(4): blk = rc.CorrBlock(f1, f2, LV, 4)
        torch.cuda.synchronize()
        ts = []
        for trial in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): blk = rc.CorrBlock(f1, f2, LV, 4)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 10 * 1e3)
    n = H * W
    alg = B * n * sum((H >> l) * (W >> l) for l in range(LV)) * 4
    tiles = B * ((n + 127) // 128) * ((H + 7) // 8) * ((W + 31) // 32)
    print(f"B={B} {H}x{W}: build group {min(ts):8.1f} us  {alg / min(ts) / 1e3:7.0f} GB/s algorithmic  {tiles} tiles  {min(ts) * 148 / tiles:.2f} us/tile/SM")
    del blk, f1, f2
    torch.cuda.empty_cache()
