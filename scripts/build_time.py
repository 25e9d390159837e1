import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, H, W = 8, 55, 128
LV = int(os.environ.get("LEVELS", "4"))
for D in [int(x) for x in (sys.argv[1:] or ["256"])]:
    f1 = torch.randn((B, D, H, W), device=dev); f2 = torch.randn((B, D, H, W), device=dev)
    with torch.no_grad():
        for _ in range(5): blk = rc.CorrBlock(f1, f2, LV, 4)
        torch.cuda.synchronize()
        ts = []
        for trial in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): blk = rc.CorrBlock(f1, f2, LV, 4)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 20 * 1e3)
        print(f"L={LV} D={D}: build group (transpose + build_tc) {min(ts):.1f} us (trials {' '.join('%.0f' % t for t in ts)})")
