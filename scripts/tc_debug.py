"""GPU-side diagnostic for the tcgen05 build: compares build_mode='tf32' with the exact fp32 build on a
ladder of shapes and prints where (which level / tile / row / column) they disagree."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import optical_flow_dnn_b200 as rc

dev = torch.device("cuda:0")
shapes = [
    (1, 32, 8, 32, 1), (1, 32, 8, 32, 3), (1, 256, 8, 32, 3), (1, 64, 16, 64, 4), (2, 128, 19, 22, 4),
    (1, 24, 16, 20, 4), (2, 256, 46, 62, 4), (1, 256, 55, 128, 4), (3, 256, 47, 156, 4),
]
ok_all = True
for (B, D, H, W, L) in shapes:
    g = torch.Generator(device=dev).manual_seed(B * 1000 + D + H + W)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev)
    f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ref = rc.CorrBlock(f1, f2, L, 4 if min(H >> (L - 1), W >> (L - 1)) >= 2 else 1, build_mode="fp32")
    torch.cuda.synchronize()
    try:
        t0 = time.time()
        got = rc.CorrBlock(f1, f2, L, ref.radius, build_mode="tf32")
        torch.cuda.synchronize()
        dt = time.time() - t0
    except Exception as e:
        print(f"shape {(B, D, H, W, L)}: EXCEPTION {e}")
        ok_all = False
        break
    N = H * W
    line = f"shape {(B, D, H, W, L)} ({dt*1e3:.1f} ms):"
    for l in range(L):
        a, b = got.corr_pyramid[l], ref.corr_pyramid[l]
        scale = float(b.abs().max())
        err = (a - b).abs()
        rel = float(err.max()) / scale
        bad = err > 2e-3 * scale
        line += f"  L{l} rel {rel:.2e} bad {int(bad.sum())}/{bad.numel()}"
        if rel > 1e-3:
            ok_all = False
            idx = bad.nonzero()[:8].tolist()
            hl, wl = b.shape[-2:]
            print(line)
            print(f"   level {l}: first bad (bn, _, h, w): {idx}")
            bn = bad.reshape(B, N, hl, wl)
            print("   bad per batch:", bn.sum(dim=(1, 2, 3)).tolist())
            print("   bad per m%128 (first 16 lanes):", bn.sum(dim=(0, 2, 3)).reshape(-1)[:N - N % 128 or N].reshape(-1, min(128, N))[:, :16].sum(0).tolist() if N >= 128 else "n/a")
            print("   bad per row h:", bn.sum(dim=(0, 1, 3)).tolist())
            print("   bad per col w:", bn.sum(dim=(0, 1, 2)).tolist())
            print("   sample got/ref:", a.reshape(-1)[bad.reshape(-1)][:6].tolist(), b.reshape(-1)[bad.reshape(-1)][:6].tolist())
            line = "   ..."
    print(line)
print("TC_DEBUG", "PASS" if ok_all else "FAIL")
