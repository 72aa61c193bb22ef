"""AtomMessagePassing fwd+bwd step time (bf16 tier): tensor-core path vs the f32-accurate SIMT kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemprop_b200.data import BatchMolGraph, make_molecules
from chemprop_b200.nn import AtomMessagePassing, MeanAggregation

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda")
bmg = BatchMolGraph(make_molecules(n, seed=1, mean_atoms=25.0)); bmg.to(dev)
agg = MeanAggregation()
for fused in (True, False):
    torch.manual_seed(0)
    mp = AtomMessagePassing(d_h=300, depth=3, precision="bf16").to(dev)
    mp.fused = fused
    params = list(mp.parameters())

    def step():
        bmg._layout = None
        for p in params: p.grad = None
        agg(mp(bmg), bmg.batch).float().square().mean().backward()

    for _ in range(3): step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"AtomMP bf16 tier, {n} mols, tensor-core path={fused}: {ms:.2f} ms/step = {n / ms * 1e3 / 1e6:.2f} M mol/s")
