"""Phase timeline (block 0) of the fused depth step via dmpnn_set_trace_buffer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemprop_b200 import _lib
from chemprop_b200.data import BatchMolGraph, make_molecules
from chemprop_b200.engine import bond_step_fused, get_layout, pack_weight_bf16, pad_hidden
h = 300
bmg = BatchMolGraph(make_molecules(10000, seed=1)); bmg.to("cuda")
lay = get_layout(bmg); hp = pad_hidden(h)
H0 = torch.zeros(lay.E, hp, dtype=torch.bfloat16, device="cuda"); H0[:, :h] = torch.randn(lay.E, h, device="cuda").bfloat16()
Hp = torch.relu(H0).clone(); Hn = torch.zeros_like(H0)
Wpk = pack_weight_bf16(torch.randn(h, h, device="cuda") / h ** 0.5)
for _ in range(3): bond_step_fused(Hp, H0, Hn, h, Wpk, None, lay, _lib.ACT_RELU, 0.0, False)
NT = 28
tr = torch.zeros(NT * 16, dtype=torch.int64, device="cuda")
lib = _lib.load(); lib.dmpnn_set_trace_buffer(tr.data_ptr(), NT)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
bond_step_fused(Hp, H0, Hn, h, Wpk, None, lay, _lib.ACT_RELU, 0.0, False)
e1.record()
torch.cuda.synchronize(); lib.dmpnn_set_trace_buffer(None, 0)
print("block", os.environ.get("DMPNN_TRACE_BLOCK", "0"), "kernel ms", e0.elapsed_time(e1), "tiles", lay.n_tiles)
t = tr.cpu().view(NT, 16).numpy().astype("float64")
t0 = t[t > 0].min()
names = ["A_issue", "S_start", "S_done", "MMA_ready", "MMA_c0go", "MMA_issued", "E0_start", "E1_start", "E0_end", "E1_end", "H0_issue", "s2_wait", "s2_full", "s2_done", "s2_bar", "s2_out"]
print("tile " + " ".join(f"{n:>10s}" for n in names[:16]))
for i in range(NT):
    print(f"{i:4d} " + " ".join(f"{(t[i, k] - t0) / 1000 if t[i, k] > 0 else float('nan'):10.2f}" for k in range(16)))
