"""Data parallelism for the path: molecules are independent, so ranks shard the
batch with no data-path collective; only parameter gradients are exchanged -- one flat f32
bucket, one all-reduce per step (NCCL on GPUs; gloo in the CPU tests).  This replaces the
DistributedDataParallel wrapper Lightning puts around the reference (chemprop/cli/train.py:1930-1939).
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

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def allreduce_(self):
        """grad <- mean over ranks (in place).  Parameters without a grad contribute zeros."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
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
