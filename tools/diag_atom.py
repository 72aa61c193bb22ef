import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemprop_b200.data import BatchMolGraph, make_molecules
from chemprop_b200.nn import AtomMessagePassing, MeanAggregation
from oracle import restatement as R
for d_h, depth, bias in ((128, 4, True), (300, 3, False), (128, 2, True), (128, 4, False)):
    torch.manual_seed(0)
    mgs = make_molecules(800, seed=5)
    bmg = BatchMolGraph(mgs)
    mp = AtomMessagePassing(d_h=d_h, depth=depth, bias=bias, precision="bf16")
    P = {k: v.detach().double().requires_grad_(True) for k, v in mp.state_dict().items()}
    H_ref = R.message_passing_forward("atom", bmg.V.double(), bmg.E.double(), bmg.edge_index, bmg.rev_edge_index,
                                      P["W_i.weight"], P.get("W_i.bias"), P["W_h.weight"], P.get("W_h.bias"),
                                      P["W_o.weight"], P["W_o.bias"], depth)
    a_ref = R.aggregate(H_ref, bmg.batch, "mean"); a_ref.square().sum().backward()
    mp = mp.cuda(); bmg.to("cuda")
    for fused in (True, False):
        mp.fused = fused
        for p in mp.parameters(): p.grad = None
        bmg._layout = None
        H = mp(bmg); a = MeanAggregation()(H, bmg.batch); a.float().square().sum().backward()
        eH = (H.detach().double().cpu() - H_ref.detach()).abs().max().item()
        out = [f"h={d_h} d={depth} bias={bias} tc={fused}: |H|max {H_ref.abs().max().item():.2f} errH {eH:.4f}"]
        for k, p in mp.named_parameters():
            ref = P[k].grad; scale = max(1e-6, ref.abs().max().item())
            out.append(f"{k} {(p.grad.double().cpu() - ref).abs().max().item() / scale:.4f}")
        print(" | ".join(out))
