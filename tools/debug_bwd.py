import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemprop_b200 import _lib
from chemprop_b200.data import BatchMolGraph, make_molecules
from chemprop_b200.engine import bond_step_bwd_fused, get_layout, pack_weight_bf16, pad_hidden
h = 64
bmg = BatchMolGraph(make_molecules(40, seed=2, shuffle_edges=True)); bmg.to("cuda"); lay = get_layout(bmg); hp = pad_hidden(h)
E = lay.E
dZ = torch.zeros(E, hp, dtype=torch.bfloat16, device="cuda")
dZ[:, :h] = torch.arange(E, device="cuda").float()[:, None].expand(E, h).bfloat16() % 64   # row id in every column
W = torch.eye(h, device="cuda")
out = torch.zeros_like(dZ)
bond_step_bwd_fused(dZ, None, out, h, pack_weight_bf16(W.t().contiguous()), lay, _lib.ACT_RELU, 0.0)
torch.cuda.synchronize()
rev, dst = lay.rev_row.long(), lay.dst_row.long()
X = dZ[:, :h].float()[rev]
A = torch.zeros(lay.V, h, device="cuda").index_add_(0, dst, X)
G = (A[dst] - X)
err = (out[:, :h].float() - G).abs()
print("E", E, "tiles", lay.n_tiles, "max err", err.max().item(), "rows wrong", (err.max(1).values > 0.5).sum().item())
bad = (err.max(1).values > 0.5).nonzero().flatten()[:10].tolist()
for r in bad:
    print("row", r, "got", out[r, 0].item(), "want", G[r, 0].item(), "dst", dst[r].item(), "rev", rev[r].item(), "tile_row_ptr", lay.tile_row_ptr[:lay.n_tiles+1].tolist()[:6])
# also variant: what if kernel computed forward-style (no rev on read)?
X2 = dZ[:, :h].float(); A2 = torch.zeros(lay.V, h, device="cuda").index_add_(0, dst, X2); G2 = A2[dst] - X2
print("matches forward-style gather:", (out[:, :h].float() - G2).abs().max().item(), " forward-style permuted out:", (out[:, :h].float()[rev] - G2).abs().max().item())
print("---- random W, copy / mask")
torch.manual_seed(0)
for hh in (64, 300):
    hp = pad_hidden(hh)
    dZ = torch.zeros(E, hp, dtype=torch.bfloat16, device="cuda"); dZ[:, :hh] = torch.randn(E, hh, device="cuda").bfloat16()
    W = torch.randn(hh, hh, device="cuda") / hh ** 0.5
    Y = torch.zeros(E, hp, dtype=torch.bfloat16, device="cuda"); Y[:, :hh] = torch.relu(torch.randn(E, hh, device="cuda")).bfloat16()
    X = dZ[:, :hh].float()[rev]; A = torch.zeros(lay.V, hh, device="cuda").index_add_(0, dst, X)
    G = (A[dst] - X).bfloat16().float() @ W.bfloat16().float()
    Wpk = pack_weight_bf16(W.t().contiguous())
    for masked in (False, True):
        out = torch.zeros(E, hp, dtype=torch.bfloat16, device="cuda")
        bond_step_bwd_fused(dZ, Y if masked else None, out, hh, Wpk, lay, _lib.ACT_RELU, 0.0)
        torch.cuda.synchronize()
        ref = G * (Y[:, :hh].float() > 0).float() if masked else G
        ref_wrongT = ((A[dst] - X).bfloat16().float() @ W.bfloat16().float().t())
        print(f"h={hh} masked={masked}: err {(out[:, :hh].float() - ref).abs().max().item():.4f}  (vs W^T: {(out[:, :hh].float() - (ref_wrongT * ((Y[:, :hh].float() > 0).float() if masked else 1))).abs().max().item():.4f}) |ref| {ref.abs().max().item():.2f}")
print("---- identity W, random dZ (column-varying)")
hh = 64; hp = pad_hidden(hh)
dZ = torch.zeros(E, hp, dtype=torch.bfloat16, device="cuda"); dZ[:, :hh] = torch.randn(E, hh, device="cuda").bfloat16()
X = dZ[:, :hh].float()[rev]; A = torch.zeros(lay.V, hh, device="cuda").index_add_(0, dst, X); G = (A[dst] - X)
out = torch.zeros(E, hp, dtype=torch.bfloat16, device="cuda")
bond_step_bwd_fused(dZ, None, out, hh, pack_weight_bf16(torch.eye(hh, device="cuda")), lay, _lib.ACT_RELU, 0.0)
torch.cuda.synchronize()
e = (out[:, :hh].float() - G.bfloat16().float()).abs()
print("err", e.max().item(), "bad rows", (e.max(1).values > 0.05).sum().item(), "bad cols", (e.max(0).values > 0.05).sum().item())
print("---- scaled identity 2I")
out = torch.zeros(E, hp, dtype=torch.bfloat16, device="cuda")
bond_step_bwd_fused(dZ, None, out, hh, pack_weight_bf16(2 * torch.eye(hh, device="cuda")), lay, _lib.ACT_RELU, 0.0)
torch.cuda.synchronize()
print("err", (out[:, :hh].float() - (2 * G).bfloat16().float()).abs().max().item())
print("---- permutation W (shift by one)")
P = torch.roll(torch.eye(hh, device="cuda"), 1, dims=1)
out = torch.zeros(E, hp, dtype=torch.bfloat16, device="cuda")
bond_step_bwd_fused(dZ, None, out, hh, pack_weight_bf16(P.t().contiguous()), lay, _lib.ACT_RELU, 0.0)
torch.cuda.synchronize()
ref = G.bfloat16().float() @ P
print("err", (out[:, :hh].float() - ref).abs().max().item(), " vs P^T:", (out[:, :hh].float() - G.bfloat16().float() @ P.t()).abs().max().item())
