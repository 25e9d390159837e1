"""Config 3 (training step, forward + backward through 12 lookups and the build) as ONE CUDA graph: device time
without the Python / autograd host overhead that dominates the eager measurement in bench_configs.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W, L, r, iters = 10, 256, 46, 62, 4, 4, 12
g = torch.Generator(device=dev).manual_seed(1234)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
coords = [(grid + flow).contiguous()]
for _ in range(iters - 1): coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
ws = [torch.randn((B, L * (2 * r + 1) ** 2, H, W), device=dev) for _ in range(iters)]
a = f1.clone().requires_grad_(); b = f2.clone().requires_grad_()
def step():
    a.grad = None; b.grad = None
    blk = rc.CorrBlock(a, b, L, r)
    if os.environ.get("TRAIN_LOSS_SUM") == "1":  # round-1 harness: 12 x (mul + sum) and their backward in torch
        loss = 0
        for c, w in zip(coords, ws): loss = loss + (blk(c) * w).sum()
        loss.backward()
    else:  # out.backward(randn) per iteration, accumulated -- only this library's kernels in the step
        torch.autograd.backward([blk(c) for c in coords], ws)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    step()
gr.replay(); torch.cuda.synchronize()
ref1 = a.grad.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): gr.replay()
e1.record(); torch.cuda.synchronize()
assert torch.equal(ref1, a.grad)
print(f"training step (B={B}, {H}x{W}, {iters} it, fwd+bwd) as a CUDA graph: {e0.elapsed_time(e1)/20:.3f} ms / step, "
      f"{B/(e0.elapsed_time(e1)/20*1e-3):.0f} pairs/s  lookup_bwd={os.environ.get('RAFTCORR_LOOKUP_BWD','batched')}")
