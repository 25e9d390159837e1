"""Where the HOST time of the eager config-3 training step goes (cProfile + allocator counters)."""
import cProfile, os, pstats, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
B, D, H, W, L, r, iters = 10, 256, 46, 62, 4, 4, 12
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
coords = [(grid + 2 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous() for _ in range(iters)]
ws = [torch.randn((B, L * 81, H, W), device=dev) for _ in range(iters)]
def step():
    a = f1.clone().requires_grad_(); b = f2.clone().requires_grad_()
    blk = rc.CorrBlock(a, b, L, r)
    torch.autograd.backward([blk(c) for c in coords], ws)
for _ in range(5): step()
torch.cuda.synchronize()
s0 = torch.cuda.memory_stats()
t0 = time.perf_counter()
for _ in range(20): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
s1 = torch.cuda.memory_stats()
print(f"20 steps: host {1e3*(t1-t0)/20:.3f} ms/step, until done {1e3*(t2-t0)/20:.3f} ms/step; cudaMalloc calls {s1['num_device_alloc']-s0['num_device_alloc']}, frees {s1['num_device_free']-s0['num_device_free']}, alloc retries {s1['num_alloc_retries']-s0['num_alloc_retries']}")
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable(); torch.cuda.synchronize()
out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(22); print(out.getvalue()[:6000])
