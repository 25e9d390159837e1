"""How much host time does one lookup call cost? (GPU box)"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
from optical_flow_dnn_b200 import _lib
dev = torch.device("cuda:0")
B, D, H, W = 1, 256, 16, 24   # tiny: the GPU is never the bottleneck, we time the enqueue path
f1 = torch.randn(B, D, H, W, device=dev); f2 = torch.randn(B, D, H, W, device=dev)
co = torch.rand(B, 2, H, W, device=dev) * 10
for mode in ("tma",):
    os.environ["RAFTCORR_LOOKUP"] = mode
    blk = rc.CorrBlock(f1, f2, 2, 4)
    for _ in range(50): blk(co)
    torch.cuda.synchronize()
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n): blk(co)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{mode}: CorrBlock.__call__ host {1e6*(t1-t0)/n:.1f} us/call (drain {1e3*(t2-t1):.1f} ms)")
    L = _lib.lib(); out = torch.empty((B, 2 * 81, H, W), device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    pp, pc, po = ctypes.c_void_p(blk._pyr.data_ptr()), ctypes.c_void_p(co.data_ptr()), ctypes.c_void_p(out.data_ptr())
    t0 = time.perf_counter()
    for _ in range(n): L.raftcorr_lookup_fwd_planned(blk._plan, pc, po, 0, stream)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{mode}: raw planned C call host {1e6*(t1-t0)/n:.1f} us/call")
    t0 = time.perf_counter()
    for _ in range(n): L.raftcorr_lookup_fwd(pp, ctypes.byref(blk._layout), pc, 4, po, 0, stream)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{mode}: raw unplanned C call host {1e6*(t1-t0)/n:.1f} us/call")
t0 = time.perf_counter()
for _ in range(3000): torch.empty((B, 162, H, W), device=dev)
print(f"torch.empty {1e6*(time.perf_counter()-t0)/3000:.1f} us")
t0 = time.perf_counter()
for _ in range(300): rc.CorrBlock(f1, f2, 2, 4)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"CorrBlock ctor host {1e6*(t1-t0)/300:.1f} us")
