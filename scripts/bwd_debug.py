import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optical_flow_dnn_b200 import _lib
dev = torch.device("cuda:0")
L = _lib.lib()
def P(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
for (B, D, H, W, LV) in [(1, 16, 9, 12, 1), (1, 32, 8, 32, 1), (2, 128, 19, 22, 3), (1, 256, 46, 62, 4), (2, 256, 55, 128, 4)]:
    lay = _lib.make_layout(B, H, W, LV)
    g = torch.Generator(device=dev).manual_seed(1)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    dp0 = torch.randn(lay.total_floats, generator=g, device=dev)
    res = {}
    for mode in (0, 1):
        dp = dp0.clone()
        df1 = torch.full((B, D, H, W), 7.0, device=dev); df2 = torch.full((B, D, H, W), 7.0, device=dev)
        wsb = L.raftcorr_build_bwd_workspace_bytes(B, D, H, W, mode)
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.raftcorr_build_bwd(P(dp), ctypes.byref(lay), P(f1), P(f2), D, P(df1), P(df2), mode, P(ws), wsb, 0, st)
        torch.cuda.synchronize()
        if rc: print("error", L.raftcorr_last_error()); break
        res[mode] = (df1, df2)
    for i, name in enumerate(("df1", "df2")):
        a, b = res[1][i], res[0][i]
        err = (a - b).abs().max().item() / b.abs().max().item()
        print(f"shape {(B,D,H,W,LV)} {name}: rel err {err:.2e}  tc: min {a.min().item():.3f} max {a.max().item():.3f} n7={(a==7).sum().item()} n0={(a==0).sum().item()} of {a.numel()}")
        if err > 1e-2:
            bad = ((a - b).abs() > 1e-2 * b.abs().max())
            print("   bad per d (first 16):", bad.sum(dim=(0, 2, 3))[:16].tolist())
            print("   bad per h:", bad.sum(dim=(0, 1, 3)).tolist())
            print("   bad per w:", bad.sum(dim=(0, 1, 2)).tolist())
            print("   ratio a/b sample:", (a / b).flatten()[:8].tolist())
