import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W = 8, 256, 55, 128
f1 = torch.randn((B, D, H, W), device=dev); f2 = torch.randn((B, D, H, W), device=dev)
with torch.no_grad():
    for _ in range(5): blk = rc.CorrBlock(f1, f2, 4, 4)
    torch.cuda.synchronize()
    for trial in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): blk = rc.CorrBlock(f1, f2, 4, 4)
        e1.record(); torch.cuda.synchronize()
        print(f"build (2 transposes + build_tc) {e0.elapsed_time(e1)/50*1e3:.1f} us")
