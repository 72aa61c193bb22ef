"""GPU timing of the elementwise / segment kernels at C2 size: 8-byte (C=300) vs 16-byte (C=304) vector paths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemprop_b200 import _lib
from chemprop_b200.data import BatchMolGraph, make_molecules
from chemprop_b200.engine import act_bwd, bond_message, bond_message_bwd_masked, get_layout, sum_act_bwd
bmg = BatchMolGraph(make_molecules(10000, seed=1)); bmg.to("cuda"); lay = get_layout(bmg)
E = lay.E
mk = lambda: torch.randn(E, 320, device="cuda").bfloat16()
X, Y, O, Z2 = mk(), mk(), mk(), mk()
dMv = torch.randn(lay.V, 320, device="cuda").bfloat16()
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); ts = []
    for _ in range(8):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3
for C in (300, 304):
    print(f"C={C}: bond_message {t(lambda: bond_message(X, lay, C, O)):.0f} us | permuted {t(lambda: bond_message(X, lay, C, O, permute_on_read=True)):.0f} us | "
          f"masked {t(lambda: bond_message_bwd_masked(X, Y, lay, C, O, act=_lib.ACT_RELU)):.0f} us | "
          f"act_bwd(gather) {t(lambda: act_bwd(dMv, Y, E, C, act=_lib.ACT_RELU, gidx=lay.dst_row, dZ=O)):.0f} us | "
          f"sum_act_bwd {t(lambda: sum_act_bwd([X, Z2], Y, O, O.new_empty(E, 320), E, C, act=_lib.ACT_RELU)):.0f} us")
