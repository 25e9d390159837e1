"""BASELINE.json configs[2] on N GPUs: training step, B=10 pairs per GPU at 368x496 (fmap 46x62, D=256, r=4, 12
iterations), forward + backward through every lookup and the volume, then ONE NCCL all-reduce (sum) of the parameter
gradients over NVLink (21.0 MB fp32 for RAFT-basic, of which the feature encoder's share is 4.3 MB).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/train_step_ddp.py

The encoders / update block are the caller's and are not rebuilt here: the feature maps come from the fnet output head
(the final 1x1 convolution 128 -> 256, raft/core/extractor.py:172) applied to seeded activations, so that real
parameter gradients flow out of the correlation backward; they sit at the front of a flat 21 MB gradient bucket (the
rest stands in for the other parameters) that is all-reduced asynchronously right after backward, the way DDP does.
Pairs are sharded by rank, no collective on the correlation path itself.  Prints one JSON line (rank 0): step time as
the max over ranks, with and without the all-reduce, and a cross-rank equality check of the reduced gradients."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import optical_flow_dnn_b200 as rc
from optical_flow_dnn_b200.ddp import GradBucket

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
B, Cin, D, H, W, L, r, iters = 10, 128, 256, 46, 62, 4, 4, 12
g = torch.Generator(device=dev).manual_seed(1234 + rank)       # every rank has its own pairs
gw = torch.Generator(device=dev).manual_seed(99)               # ... and the same parameters
x1 = torch.randn((B, Cin, H, W), generator=g, device=dev); x2 = torch.randn((B, Cin, H, W), generator=g, device=dev)
head_w = (torch.randn((D, Cin, 1, 1), generator=gw, device=dev) / Cin ** 0.5).requires_grad_()
head_b = torch.zeros((D,), device=dev).requires_grad_()
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
coords = [(grid + flow).contiguous()]
for _ in range(iters - 1): coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
ws = [torch.randn((B, L * (2 * r + 1) ** 2, H, W), generator=g, device=dev) for _ in range(iters)]
BUCKET = 21_000_000 // 4                                       # RAFT-basic: 5.25 M fp32 parameters
n_head = head_w.numel() + head_b.numel()
rest = torch.zeros(BUCKET - n_head, device=dev, requires_grad=True)   # stands in for the other parameters' gradients
gb = GradBucket([head_w, head_b, rest])                        # .grad tensors are views of ONE flat buffer
bucket = gb.flat
FUSED_HEAD = os.environ.get("TRAIN_FUSED_HEAD", "1") != "0"    # CorrBlock.from_encoder_head vs einsum head + CorrBlock
x12 = torch.cat([x1, x2]).contiguous()

def step(allreduce):
    gb.zero()
    if FUSED_HEAD:   # conv2 head fused into the build; gradients flow back through it
        blk = rc.CorrBlock.from_encoder_head(x12, head_w, head_b, num_levels=L, radius=r)
    else:
        w2 = head_w.view(D, Cin)  # the 1x1 convolution as a plain product (cuBLAS), + bias
        f1 = torch.einsum("oc,bchw->bohw", w2, x1) + head_b.view(1, D, 1, 1)
        f2 = torch.einsum("oc,bchw->bohw", w2, x2) + head_b.view(1, D, 1, 1)
        blk = rc.CorrBlock(f1, f2, L, r)
    if os.environ.get("TRAIN_LOSS_SUM") == "1":   # round-1 harness: 12 x (mul + sum) and their backward in torch
        loss = 0
        for c, w in zip(coords, ws): loss = loss + (blk(c) * w).sum()
        loss.backward()          # accumulates straight into the bucket
    else:                        # out.backward(randn) per iteration, accumulated
        torch.autograd.backward([blk(c) for c in coords], ws)
    if allreduce:
        gb.all_reduce_async()   # side stream, averaged over ranks
        gb.wait()

# TRAIN_GRAPH=1: forward + backward replayed as ONE CUDA graph (1.3 ms of device work behind ~2 ms of eager Python /
# autograd launches otherwise: with a collective in the step, every rank then waits for the slowest rank's HOST each
# step); the all-reduce stays outside the graph, on the bucket's side stream.
GRAPH = os.environ.get("TRAIN_GRAPH", "0") == "1"
if GRAPH:
    eager_step = step
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3): eager_step(False)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eager_step(False)

    def step(allreduce):
        graph.replay()
        if allreduce:
            gb.all_reduce_async()
            gb.wait()

def timed(allreduce, n=10):
    for _ in range(6): step(allreduce)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step(allreduce)
    e1.record()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

ms_local = timed(False)
ms_ar = timed(True)
# all-reduce alone on the 21 MB bucket
if world > 1:
    for _ in range(3): dist.all_reduce(bucket)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_reduce(bucket)
    e1.record(); torch.cuda.synchronize()
    ar_us = e0.elapsed_time(e1) / 20 * 1e3
else:
    ar_us = 0.0
# correctness of the reduction: every rank holds the same reduced head gradient, equal to the sum of the local ones
step(False)
local = bucket[:n_head].clone()
ok = True
if world > 1:
    summed = local.clone(); dist.all_reduce(summed)
    gathered = [torch.empty_like(summed) for _ in range(world)]
    dist.all_gather(gathered, summed)
    ok = all(torch.equal(gathered[0], t_) for t_ in gathered) and bool(torch.isfinite(summed).all())
if rank == 0:
    print(json.dumps({"fused_head": FUSED_HEAD, "cuda_graph": GRAPH, "config": "3 training step, B=10 pairs/GPU, 46x62, D=256, r=4, 12 it, fwd+bwd (" + ("one CUDA graph" if GRAPH else "eager autograd") + ")", "n_gpus": world,
                      "ms_per_step_no_allreduce": ms_local, "ms_per_step_with_allreduce": ms_ar,
                      "pairs_per_s": world * B / (ms_ar * 1e-3), "allreduce_21MB_us": ar_us, "grads_identical_across_ranks": ok}))
if world > 1: dist.destroy_process_group()
