"""Where does the host time of bench.py's loader-driven (e2e) step go?  Per-phase wall-clock of the training thread over
30 steps of the C2 configuration (resident packed data set + PackedBatchLoader.stream()), and a cProfile of the same loop.
    python tools/profile_e2e.py [n_steps]"""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from chemprop_b200.data import PackedBatchLoader, PackedMolGraphDataset, make_molecules
from chemprop_b200.nn import BondMessagePassing, MeanAggregation

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mp = BondMessagePassing(precision="bf16").to(dev)
agg = MeanAggregation()
params = list(mp.parameters())
ds = PackedMolGraphDataset.from_molgraphs(make_molecules(30000, seed=1)).to(dev)
loader = PackedBatchLoader(ds, batch_size=10000, shuffle=True, seed=5, drop_last=True)
it = loader.stream()
loss_host = torch.empty(2, dtype=torch.float32).pin_memory()


def run(k, T=None):
    pending = None
    for i in range(k):
        t0 = time.perf_counter()
        for p in params:
            p.grad = None
        b = next(it)
        t1 = time.perf_counter()
        bmg = b.bmg
        bmg._layout = None
        H = mp(bmg)
        loss = agg(H, bmg.batch).float().square().mean()
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        slot = loss_host[i & 1:(i & 1) + 1]
        slot.copy_(loss.detach().float().reshape(1), non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        if pending is not None:
            pending.synchronize()
        pending = done
        t4 = time.perf_counter()
        if T is not None:
            T.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    pending.synchronize()


run(6)
torch.cuda.synchronize()
T = []
t0 = time.perf_counter()
run(n_steps, T)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n_steps
a = np.array(T) * 1e3
print(f"wall per step {wall * 1e3:.3f} ms;  host ms per step (median / mean / max):")
for name, col in zip(("next(loader)", "forward issue", "backward issue", "loss copy + wait for step i-1"), a.T):
    print(f"  {name:32s} {np.median(col):7.3f} {col.mean():7.3f} {col.max():7.3f}")
pr = cProfile.Profile()
pr.enable()
run(n_steps)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(38)
print(s.getvalue()[:6000])
