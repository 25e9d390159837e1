import os, sys
sys.path.insert(0, "/root/repo")
import torch, ctypes
import optical_flow_dnn_b200 as rc
from optical_flow_dnn_b200 import _lib
dev = torch.device("cuda:0")
B, Cin, D, H, W = 8, 128, 256, 55, 128
xs = [torch.randn((2 * B, Cin, H, W), device=dev) for _ in range(3)]
w = torch.randn((D, Cin), device=dev) / Cin ** 0.5; b = torch.randn((D,), device=dev)
lib = _lib.lib()
# time the head alone: call the alt-pyramid entry with L=1 (head + nothing else)
f1n = torch.empty((B, H, W, D), device=dev); f2p = torch.empty((B * H * W * D + 64,), device=dev)
wsb = int(lib.raftcorr_fnet_head_alt_workspace_bytes(Cin, D)); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
def run(i):
    _lib.check(lib.raftcorr_fnet_head_alt_pyramid(p(xs[i % 3]), p(w), p(b), B, Cin, D, H, W, 1, p(f1n), p(f2p), p(ws), wsb, 0, st), "head")
for i in range(5): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30): run(i)
e1.record(); torch.cuda.synchronize()
print(f"RAFTCORR_HEAD_ABL={os.environ.get('RAFTCORR_HEAD_ABL', '0')}: head (round_weight + fnet_head_kernel) {e0.elapsed_time(e1) / 30 * 1e3:.1f} us")
