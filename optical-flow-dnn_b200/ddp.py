"""One process per GPU in place of ``nn.DataParallel`` (raft/train.py:70-71,118-128).

The reference wraps RAFT in ``nn.DataParallel``: every step it re-broadcasts all parameters (21 MB for RAFT-basic) from
GPU 0 to the replicas, scatters the batch, gathers the predictions and reduces the gradients back onto GPU 0, all from
one Python process.  The correlation path itself shards by frame pair with no collective (``sharding.py``); what a
training step needs on top is exactly two things, provided here for the ``torchrun`` model:

* :func:`broadcast_parameters` -- once, at start-up (and after loading a checkpoint), instead of every step;
* :class:`GradBucket` -- ONE flat fp32 buffer that the parameters' ``.grad`` tensors are views of, so backward writes
  straight into it and a single all-reduce (NCCL over NVLink on the box, gloo in the CPU tests) averages everything:
  no per-parameter collectives, no flatten/unflatten copies.  ``all_reduce_async()`` launches on a side stream right
  after backward so the collective overlaps whatever the caller does next (gradient clipping needs ``wait()`` first);
  with equal shards per rank the averaged gradient equals the full-batch gradient DataParallel produces.

Host-side plumbing only; no kernels of this library are involved.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers (one flat broadcast per dtype)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    for ts in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            t.copy_(flat[off: off + t.numel()].view_as(t))
            off += t.numel()


class GradBucket:
    """Flat gradient buffer shared by ``params`` (those with ``requires_grad``); see the module docstring.

    >>> bucket = GradBucket(model.parameters())
    >>> loss.backward()                 # autograd accumulates into views of bucket.flat
    >>> bucket.all_reduce_async(); bucket.wait()
    >>> optimizer.step(); bucket.zero()
    """

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no parameter requires grad")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("GradBucket: parameters must share one device and dtype")
        self.group = group
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=dt, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()
        self._stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self._work = None

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def zero(self) -> None:
        """Use instead of ``optimizer.zero_grad()`` (which would drop the views with ``set_to_none=True``)."""
        self.flat.zero_()

    def all_reduce_async(self) -> None:
        """Average the bucket over the ranks; returns immediately."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat.device))  # after backward
            with torch.cuda.stream(self._stream):
                self.flat.div_(world)
                self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            self.flat.div_(world)
            self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def wait(self) -> None:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._stream)
