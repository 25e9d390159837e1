"""world_size-2 gloo tests (CPU) of the one-process-per-GPU training plumbing: parameters
broadcast once, gradients averaged through ONE flat bucket, result == the single-process full-batch step that
nn.DataParallel computes (raft/train.py:70-71,118-128).  A small convolutional stand-in plays the model: the
collective logic does not depend on what produced the gradients."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optical_flow_dnn_b200.ddp import GradBucket, broadcast_parameters
from optical_flow_dnn_b200.sharding import shard


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 2, 1))


def _batch():
    g = torch.Generator().manual_seed(7)
    return torch.randn((6, 3, 10, 12), generator=g), torch.randn((6, 2, 10, 12), generator=g)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = _model(100 + rank)              # ranks start from DIFFERENT weights ...
    broadcast_parameters(model, src=0)      # ... and end up with rank 0's
    bucket = GradBucket(model.parameters())
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    x, y = _batch()
    xs, ys = shard(x, rank, world), shard(y, rank, world)
    for _ in range(2):                      # two steps: zero() must keep the views alive
        bucket.zero()
        loss = ((model(xs) - ys) ** 2).mean()
        loss.backward()
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in bucket.params)
        bucket.all_reduce_async()
        bucket.wait()
        opt.step()
    ret[rank] = [p.detach().clone() for p in model.parameters()]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_full_batch_step():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    # single process, full batch (what DataParallel's gather + loss on GPU 0 computes)
    model = _model(100)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    x, y = _batch()
    for _ in range(2):
        opt.zero_grad()
        ((model(x) - y) ** 2).mean().backward()
        opt.step()
    for r in range(world):
        for got, want in zip(ret[r], model.parameters()):
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    for a, b in zip(ret[0], ret[1]):
        assert torch.equal(a, b)            # replicas stay bit-identical


def test_bucket_single_process_is_a_noop_collective():
    model = _model(1)
    bucket = GradBucket(model.parameters())
    assert bucket.nbytes == 4 * sum(p.numel() for p in model.parameters())
    x, y = _batch()
    ((model(x) - y) ** 2).mean().backward()
    before = bucket.flat.clone()
    bucket.all_reduce_async()
    bucket.wait()
    assert torch.equal(before, bucket.flat) and before.abs().sum() > 0
