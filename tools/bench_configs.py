"""GPU: the other BASELINE.json configurations as one JSON line each (bench.py itself stays on config 2, the one the
metric is quoted on; these are parity-test shapes, timed here for orientation):
  C3  50 k molecules, BondMessagePassing h = 600, depth 6, fp32 tier (SIMT f32 GEMMs today: the split-bf16 tensor-core path is next)
  C4  10 k ~80-atom condensed reaction graphs, d_v = 106, d_e = 28, AtomMessagePassing h = 300, depth 3, bf16 tier
  C2a config 2 with AtomMessagePassing (atom-granular tensor-core path)
Usage: python tools/bench_configs.py [C3 C4 C2a] [--steps K]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemprop_b200 import _lib
from chemprop_b200.data import BatchMolGraph, make_cgr_graphs, make_molecules, tile_packing_order_of
from chemprop_b200.nn import AtomMessagePassing, BondMessagePassing, MeanAggregation

CONFIGS = {
    "C3": dict(graphs=lambda: make_molecules(50_000, seed=1), cls=BondMessagePassing, kw=dict(d_h=600, depth=6, precision="fp32")),
    "C4": dict(graphs=lambda: make_cgr_graphs(10_000, seed=1), cls=AtomMessagePassing,
               kw=dict(d_v=106, d_e=28, d_h=300, depth=3, precision="bf16")),
    "C2a": dict(graphs=lambda: make_molecules(10_000, seed=1), cls=AtomMessagePassing, kw=dict(d_h=300, depth=3, precision="bf16")),
}


def run(name, steps):
    cfg = CONFIGS[name]
    mgs = cfg["graphs"]()
    mgs = [mgs[i] for i in tile_packing_order_of(mgs)]
    bmg = BatchMolGraph(mgs)
    n, V, E = len(bmg), bmg.V.shape[0], bmg.E.shape[0]
    bmg.to("cuda")
    torch.manual_seed(0)
    mp, agg = cfg["cls"](**cfg["kw"]).cuda(), MeanAggregation()
    params = list(mp.parameters())
    lib = _lib.load()

    def step():
        bmg._layout = None
        for p in params:
            p.grad = None
        loss = agg(mp(bmg), bmg.batch).float().square().mean()
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    l0 = lib.dmpnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": name, "module": cfg["cls"].__name__, **cfg["kw"], "molecules": n, "atoms": V, "directed_edges": E,
                      "ms_per_step": ms, "molecules_per_s": n / (ms * 1e-3), "gpu_launches_per_step": (lib.dmpnn_launch_count() - l0) / steps,
                      "loss_finite": bool(torch.isfinite(loss))}))


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in CONFIGS] or list(CONFIGS)
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 5
    for nm in names:
        run(nm, steps)
