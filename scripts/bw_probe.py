import torch, time
dev = torch.device("cuda:0")
n = 512 * 1024 * 1024  # 2 GiB fp32
x = torch.empty(n, device=dev); y = torch.empty(n, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.0)); print(f"fill  (write only): {n*4/ms/1e6:.0f} GB/s")
ms = t(lambda: x.zero_()); print(f"zero_ (memset)    : {n*4/ms/1e6:.0f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy  (read+write): {2*n*4/ms/1e6:.0f} GB/s")
ms = t(lambda: x.sum()); print(f"sum   (read only) : {n*4/ms/1e6:.0f} GB/s")
