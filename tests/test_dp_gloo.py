"""CPU tier: the N>1 path with world_size 2 on gloo -- molecule sharding and the
flat-bucket gradient all-reduce of chemprop_b200/parallel.py (NCCL on GPUs, same code)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chemprop_b200.nn import BondMessagePassing
    from chemprop_b200.parallel import FlatGradAllReducer, shard_indices

    torch.manual_seed(0)                       # identical parameters on every rank
    mp_mod = BondMessagePassing(d_h=16)
    params = list(mp_mod.parameters())
    red = FlatGradAllReducer(params)
    # rank-dependent fake gradients; one parameter deliberately without a grad on rank 1
    for i, p in enumerate(params):
        if not (rank == 1 and i == 0):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    red.allreduce_()
    exp0 = 1.0 * 1 / 2                         # (1*1 + 0) / 2 for parameter 0 (rank 1 had no grad)
    ok = abs(params[0].grad.mean().item() - exp0) < 1e-6
    for i, p in enumerate(params[1:], start=1):
        ok = ok and abs(p.grad.mean().item() - (1 + 2) * (i + 1) / 2) < 1e-6
    # attached mode: p.grad are views of the bucket, no copies; zero_() is one memset
    red2 = FlatGradAllReducer(params).attach()
    red2.zero_()
    for i, p in enumerate(params):
        p.grad.add_(float(rank + 1) * (i + 2))                      # what autograd's in-place accumulation does
        ok = ok and p.grad.data_ptr() == red2.views[i].data_ptr()
    red2.allreduce_()
    red2.wait()
    for i, p in enumerate(params):
        ok = ok and abs(p.grad.mean().item() - (1 + 2) * (i + 2) / 2) < 1e-6
    shards = list(shard_indices(11, rank, world))
    q.put((rank, ok, shards, red.nbytes))
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert sorted(res[0][2] + res[1][2]) == list(range(11)) and not set(res[0][2]) & set(res[1][2])
    assert res[0][3] == res[1][3] == 4 * sum(p.numel() for p in __import__("chemprop_b200").nn.BondMessagePassing(d_h=16).parameters())


def _loader_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np

    from chemprop_b200.data import BatchMolGraph, PackedBatchLoader, PackedMolGraphDataset, make_molecules

    mgs = make_molecules(101, seed=5, mean_atoms=8, std_atoms=3, min_atoms=1)            # every rank packs the same data set
    ds = PackedMolGraphDataset.from_molgraphs(mgs)
    loader = PackedBatchLoader(ds, batch_size=16, shuffle=True, seed=42, rank=rank, world=world)
    ok, mine = True, []
    for epoch in range(2):
        ids = []
        for b in loader:
            ref = BatchMolGraph([mgs[i] for i in b.ids])
            ok = ok and torch.equal(b.bmg.V, ref.V) and torch.equal(b.bmg.edge_index, ref.edge_index)
            ids.append(b.ids)
        mine.append(np.concatenate(ids))
    # the ranks' shards of an epoch: equal sizes, together the whole (padded) epoch
    for e in range(2):
        t = torch.from_numpy(mine[e].copy())
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        allids = torch.cat(gathered).numpy()
        ok = ok and len(t) == 51 and set(allids.tolist()) == set(range(101)) and len(allids) == 102
    ok = ok and not np.array_equal(np.sort(mine[0]), np.sort(mine[1]))                    # reshuffled between epochs
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_loader_shards_the_epoch_across_ranks_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_loader_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_single_process_is_noop():
    from chemprop_b200.parallel import FlatGradAllReducer

    w = torch.nn.Parameter(torch.ones(3))
    w.grad = torch.full((3,), 2.0)
    FlatGradAllReducer([w]).allreduce_()
    assert torch.equal(w.grad, torch.full((3,), 2.0))
