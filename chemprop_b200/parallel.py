"""Data parallelism for the path: molecules are independent, so ranks shard the
batch with no data-path collective; only parameter gradients are exchanged -- one flat f32
bucket, one all-reduce per step (NCCL on GPUs; gloo in the CPU tests).  This replaces the
DistributedDataParallel wrapper Lightning puts around the reference (chemprop/cli/train.py:1930-1939).

Two ways to use `FlatGradAllReducer`:

  * plain (`allreduce_()` after backward): gradients are copied into the bucket, summed, divided, copied back -- works with any
    training loop, on the compute stream;
  * attached (`attach()` once, `zero_()` instead of `zero_grad`, `allreduce_()` after backward, `wait()` before the optimizer):
    every `p.grad` IS a view of the bucket (DDP's gradient_as_bucket_view), so there are no copies; the mean is taken by the
    collective itself (NCCL's AVG), launched on a SIDE stream behind an event recorded after the backward, and the compute
    stream only waits for it where the gradients are consumed -- whatever the host queues in between (the next batch's
    gather, the loss read-back) overlaps the 0.9 MB all-reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> range:
    """Rank `rank` takes items rank, rank+world, ... (DistributedSampler's strided split)."""
    return range(rank, n_items, world)


class FlatGradAllReducer:
    """Averages the gradients of `params` over the process group through ONE flat bucket."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.attached = False
        self._side = None
        self._done = None

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    # ---- attached mode -------------------------------------------------------------------------------------------
    def attach(self):
        """Make every parameter's `.grad` a view of the bucket (f32 parameters).  Use `zero_()` instead of
        `zero_grad(set_to_none=True)` from then on: autograd accumulates into the views in place."""
        for p, v in zip(self.params, self.views):
            if p.dtype != torch.float32:
                raise TypeError("attach() needs f32 parameters (the bucket is f32)")
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v
        self.attached = True
        return self

    def zero_(self):
        """One memset for all gradients (attached mode)."""
        self.flat.zero_()

    def wait(self):
        """Make the current stream wait for the last all-reduce (attached mode, side stream); call before the optimizer
        step / before reading the gradients.  No-op otherwise."""
        if self._done is not None:
            torch.cuda.current_stream(self.flat.device).wait_event(self._done)
            self._done = None

    def _views_in_place(self) -> bool:
        return self.attached and all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(self.params, self.views))

    # ---- the collective ------------------------------------------------------------------------------------------
    def allreduce_(self):
        """grad <- mean over ranks (in place).  Parameters without a grad contribute zeros."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return
        if self._views_in_place():
            if self.flat.is_cuda:
                dev = self.flat.device
                if self._side is None:
                    self._side = torch.cuda.Stream(device=dev)
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(dev))          # the backward that filled the bucket
                self._side.wait_event(ready)
                with torch.cuda.stream(self._side):
                    dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
                    self._done = torch.cuda.Event()
                    self._done.record(self._side)
                self.flat.record_stream(self._side)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.div_(world)
            return
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
