import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optical_flow_dnn_b200 import _lib
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
L = _lib.lib()
def P(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
B, D, H, W, LV = 10, 256, 46, 62, 4
lay = _lib.make_layout(B, H, W, LV)
f1 = torch.randn((B, D, H, W), device=dev); f2 = torch.randn((B, D, H, W), device=dev)
dp0 = torch.randn(lay.total_floats, device=dev)
df1 = torch.empty_like(f1); df2 = torch.empty_like(f2)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for mode in (1, 0):
    wsb = L.raftcorr_build_bwd_workspace_bytes(B, D, H, W, mode)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
    for rep in range(4):
        dp = dp0.clone(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        rcode = L.raftcorr_build_bwd(P(dp), ctypes.byref(lay), P(f1), P(f2), D, P(df1), P(df2), mode, P(ws), wsb, 0, st)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"mode {mode} rep {rep}: host call {1e3*(t1-t0):.3f} ms, until done {1e3*(t2-t0):.3f} ms rc={rcode}")
# full autograd step timing
g = torch.Generator(device=dev).manual_seed(0)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
coords = [(grid + 2 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous() for _ in range(12)]
ws_ = [torch.randn((B, LV * 81, H, W), device=dev) for _ in range(12)]
for rep in range(4):
    a = f1.clone().requires_grad_(); b = f2.clone().requires_grad_()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    blk = rc.CorrBlock(a, b, LV, 4)
    loss = 0
    for c, w in zip(coords, ws_): loss = loss + (blk(c) * w).sum()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"step rep {rep}: fwd {1e3*(t1-t0):.2f} ms | backward() host {1e3*(t2-t1):.2f} ms, done {1e3*(t3-t1):.2f} ms")
