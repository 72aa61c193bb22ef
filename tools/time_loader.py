"""GPU: molecules/s of the three ways a step can get its batch -- (a) PackedBatchLoader over a host data set (gather
thread -> pinned ring -> side-stream H2D), (b) the same with the bf16 / int32 transfer copy, (c) the HBM-resident data set
(one gather launch) -- each feeding the bf16 fwd+bwd step of bench.py's workload; and the gather kernel alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chemprop_b200.data import PackedBatchLoader, PackedMolGraphDataset, make_molecules
from chemprop_b200.nn import BondMessagePassing, MeanAggregation

pool, bs = 60000, 10000
dev = torch.device("cuda")
ds = PackedMolGraphDataset.from_molgraphs(make_molecules(pool, seed=1), pin_memory=True)
torch.manual_seed(0)
mp, agg = BondMessagePassing(precision="bf16").to(dev), MeanAggregation()
params = list(mp.parameters())

def step(bmg):
    for p in params: p.grad = None
    loss = agg(mp(bmg), bmg.batch).float().square().mean()
    loss.backward()
    return loss

def run(loader, epochs=3):
    n, losses = 0, []
    for b in loader: step(b.bmg)                       # warm-up epoch
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(epochs):
        for b in loader:
            losses.append(step(b.bmg).detach()); n += len(b.ids)
    assert bool(torch.isfinite(torch.stack(losses)).all())
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0)

print("host data set -> loader (f32/int64 H2D):   %.2f M molecules/s" % (run(PackedBatchLoader(ds, bs, seed=0, device=dev)) / 1e6))
print("host data set -> loader (bf16/int32 H2D):  %.2f M molecules/s" % (run(PackedBatchLoader(ds, bs, seed=0, device=dev, transfer_dtype=torch.bfloat16)) / 1e6))
dsd = ds.to(dev)
print("resident data set -> loader:               %.2f M molecules/s" % (run(PackedBatchLoader(dsd, bs, seed=0, device=dev)) / 1e6))
ids = np.random.default_rng(0).permutation(pool)[:bs]
for _ in range(3): dsd.batch(ids)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): b = dsd.batch(ids)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
nbytes = 2 * (b.V.numel() * 4 + b.E.numel() * 4) + 40 * b.E.shape[0] + 8 * b.V.shape[0]
print("dmpnn_dataset_gather alone: %.3f ms per 10 k molecules = %.0f GB/s algorithmic" % (ms, nbytes / ms / 1e6))
