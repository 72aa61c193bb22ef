"""Diagnostic (GPU): error of the bf16 tier vs the reference goldens, per case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import build_engine_module, golden_bmg, golden_names, load_golden
from chemprop_b200.nn import MeanAggregation

for fused in (True, False):
    print("fused =", fused)
    for name in golden_names():
        g = load_golden(name)
        mp = build_engine_module(g, "cuda", "bf16", fused)
        bmg = golden_bmg(g, "cuda")
        V_d = torch.from_numpy(g["V_d"]).cuda() if "V_d" in g else None
        H = mp(bmg, V_d)
        a = MeanAggregation()(H, bmg.batch)
        (a.float() * torch.from_numpy(g["G"]).cuda()).sum().backward()
        eh = np.abs(H.detach().float().cpu().numpy() - g["H_v"]).max()
        ea = np.abs(a.detach().float().cpu().numpy() - g["agg_mean"]).max()
        gr = []
        for k, v in g.items():
            if k.startswith("grad."):
                got = dict(mp.named_parameters())[k[5:]].grad.float().cpu().numpy()
                gr.append(f"{k[5:]}:{np.abs(got - v).max() / max(1e-12, np.abs(v).max()):.3f}/{np.linalg.norm(got - v) / max(1e-12, np.linalg.norm(v)):.3f}")
        print(f"  {name:22s} |H|max={np.abs(g['H_v']).max():.2f} errH={eh:.2e} errAgg={ea:.2e} gradrel " + " ".join(gr))
