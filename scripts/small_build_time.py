"""Build + lookup latency at small sizes (config 1: B=1, D=128, 46x62, r=3; config 2 at B=1), per build mode, interleaved trials."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
def t(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, D, H, W, r) in [(1, 128, 46, 62, 3), (1, 256, 55, 128, 4), (2, 256, 55, 128, 4), (10, 256, 46, 62, 4)]:
    f1 = torch.randn((B, D, H, W), device=dev); f2 = torch.randn((B, D, H, W), device=dev)
    res = {}
    with torch.no_grad():
        for trial in range(3):
            for mode in ("f16", "tf32", "f16x3"):
                res.setdefault(mode, []).append(t(lambda: rc.CorrBlock(f1, f2, 4, r, build_mode=mode)))
    print(f"B={B} D={D} {H}x{W}: build us  " + "  ".join(f"{m}: {min(v):.1f} ({' '.join('%.0f' % x for x in v)})" for m, v in res.items()), flush=True)
