"""world_size-2 gloo tests (CPU) of the N>1 host logic: pair sharding, gather order, max-over-ranks timing.
The CPU oracle stands in for the kernels here (tests may use it as the checker/compute stand-in)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optical_flow_dnn_b200.sharding import shard_range


def test_shard_range_partitions_exactly():
    for n in [0, 1, 2, 7, 8, 9, 16]:
        for world in [1, 2, 3, 4, 8]:
            seen = []
            for r in range(world):
                seen += list(shard_range(n, r, world))
            assert seen == list(range(n))
            # torch.chunk / DataParallel scatter semantics
            chunks = torch.arange(n).chunk(world) if n else []
            for r in range(world):
                want = chunks[r].tolist() if r < len(chunks) else []
                assert list(shard_range(n, r, world)) == want
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, ret):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import corr_oracle as O
    from optical_flow_dnn_b200.sharding import gather_pairs, max_over_ranks, shard

    rng = np.random.default_rng(99)  # same data on every rank, like a scattered global batch
    B, D, H, W, L, r = n_pairs, 16, 9, 12, 3, 2
    f1 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
    co = O.coords_grid(B, H, W) + rng.normal(0, 2, (B, 2, H, W)).astype(np.float32)
    full = O.lookup(O.corr_pyramid(f1, f2, L), co, r)
    a, b, c = (shard(torch.from_numpy(x), rank, world).numpy() for x in (f1, f2, co))
    if a.shape[0]:
        local = O.lookup(O.corr_pyramid(a, b, L), c, r)
    else:
        local = np.zeros((0,) + full.shape[1:], np.float32)
    got = gather_pairs(torch.from_numpy(local), n_pairs).numpy()
    # batch sharding is exact: no cross-pair arithmetic anywhere on the path
    ok = np.array_equal(got, full)
    t = max_over_ranks(1.0 + rank)
    dist.barrier()
    ret[rank] = (bool(ok), t)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [4, 3, 1])
def test_two_rank_sharded_lookup_matches_single_process(n_pairs):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))
    assert all(ret[r][1] == 2.0 for r in range(world))  # max over ranks
