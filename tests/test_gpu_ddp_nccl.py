"""NCCL test of the one-process-per-GPU training plumbing (raft/train.py:70-71,118-128) with THIS
library's CorrBlock in the loop: two ranks on two GPUs, pairs sharded by rank (no collective on the correlation path),
parameter gradients averaged through ONE flat bucket over NVLink -- equal to the single-GPU full-batch step that
nn.DataParallel computes.  Skipped on a one-GPU box (the CPU twin is tests/test_ddp_gloo.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
B, CIN, D, H, W, L, R, ITERS = 4, 32, 64, 24, 40, 4, 4, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(dev):
    g = torch.Generator().manual_seed(11)
    x = torch.randn((2 * B, CIN, H, W), generator=g)           # layer3 output of both frames, frame-1 images first
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1)
    coords = [grid + 2.0 * torch.randn((B, 2, H, W), generator=g) for _ in range(ITERS)]
    ws = [torch.randn((B, L * (2 * R + 1) ** 2, H, W), generator=g) for _ in range(ITERS)]
    return x.to(dev), [c.to(dev) for c in coords], [w.to(dev) for w in ws]


def _head(seed, dev):
    torch.manual_seed(seed)
    return torch.nn.Conv2d(CIN, D, 1).to(dev)                  # the encoders' conv2 output head (extractor.py:172)


def _loss(head, x1, x2, coords, ws, n_total):
    import optical_flow_dnn_b200 as rc

    blk = rc.CorrBlock(head(x1), head(x2), num_levels=L, radius=R)
    loss = 0
    for c, w in zip(coords, ws):
        loss = loss + (blk(c) * w).sum()
    return loss / n_total


def _worker(rank, world, port, ret):
    import optical_flow_dnn_b200  # noqa: F401
    from optical_flow_dnn_b200.ddp import GradBucket, broadcast_parameters
    from optical_flow_dnn_b200.sharding import shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    head = _head(100 + rank, dev)            # ranks start from DIFFERENT weights ...
    broadcast_parameters(head, src=0)        # ... and end up with rank 0's
    bucket = GradBucket(head.parameters())
    x, coords, ws = _problem(dev)
    x1, x2 = shard(x[:B], rank, world), shard(x[B:], rank, world)
    cs, wsh = [shard(c, rank, world) for c in coords], [shard(w, rank, world) for w in ws]
    for _ in range(2):                       # twice: zero() must keep the views, the accumulator must reset
        bucket.zero()
        # each rank's loss is normalised by ITS pairs; the bucket averages over ranks -> the full-batch mean
        _loss(head, x1, x2, cs, wsh, n_total=x1.shape[0]).backward()
        bucket.all_reduce_async()
        bucket.wait()
    torch.cuda.synchronize()
    ret[rank] = [p.grad.detach().cpu().clone() for p in head.parameters()]
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_step_equals_full_batch_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (tests/test_ddp_gloo.py covers the host logic on CPU)")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    head = _head(100, dev)
    x, coords, ws = _problem(dev)
    _loss(head, x[:B], x[B:], coords, ws, n_total=B).backward()
    for r in range(world):
        for got, p in zip(ret[r], head.parameters()):
            want = p.grad.cpu()
            assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-7
    for a, b in zip(ret[0], ret[1]):
        assert torch.equal(a, b)             # replicas hold bit-identical reduced gradients
