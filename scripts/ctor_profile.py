import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W = 1, 256, 16, 24
f1 = torch.randn(B, D, H, W, device=dev); f2 = torch.randn(B, D, H, W, device=dev)
for mode in ("f16", "tf32"):
    for _ in range(50): rc.CorrBlock(f1, f2, 2, 4, build_mode=mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500): rc.CorrBlock(f1, f2, 2, 4, build_mode=mode)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{mode}: ctor host {1e6*(t1-t0)/500:.1f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(500): rc.CorrBlock(f1, f2, 2, 4, build_mode="f16")
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
