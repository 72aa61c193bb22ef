"""GPU tier: the engine inside a training loop -- mirrors the reference's "can it overfit" integration tests
(tests/integration/test_regression_mol.py:56-89: 50 epochs, mse <= 0.05) with the same module combinations,
a linear head and Adam, on synthetic molecules whose target is a simple graph statistic."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _data(n=96, seed=0):
    from chemprop_b200.data import BatchMolGraph, make_molecules

    mgs = make_molecules(n, seed=seed, mean_atoms=12, std_atoms=4)
    y = np.array([[mg.V.shape[0] / 12.0 - 1.0 + 0.1 * mg.E[:, 0].sum() / max(1, mg.E.shape[0])] for mg in mgs], dtype=np.float32)
    bmg = BatchMolGraph(mgs)
    bmg.to("cuda")
    return bmg, torch.from_numpy(y).cuda()


@pytest.mark.parametrize("mp_cls,agg_cls,precision,act", [
    ("bond", "mean", "fp32", "relu"), ("atom", "sum", "fp32", "relu"), ("bond", "norm", "fp32", "relu"),
    ("bond", "mean", "bf16", "relu"), ("bond", "mean", "bf16", "tanh"),
])
def test_overfit(mp_cls, agg_cls, precision, act):
    from chemprop_b200.nn import (AtomMessagePassing, BondMessagePassing, MeanAggregation, NormAggregation,
                                  SumAggregation)

    torch.manual_seed(0)
    bmg, y = _data()
    mp = {"bond": BondMessagePassing, "atom": AtomMessagePassing}[mp_cls](d_h=64, depth=3, activation=act,
                                                                          precision=precision).cuda()
    agg = {"mean": MeanAggregation, "sum": SumAggregation, "norm": NormAggregation}[agg_cls]()
    head = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 1)).cuda()
    opt = torch.optim.Adam(list(mp.parameters()) + list(head.parameters()), lr=3e-3)
    losses = []
    for _ in range(150):
        opt.zero_grad()
        pred = head(agg(mp(bmg), bmg.batch).float())
        loss = torch.nn.functional.mse_loss(pred, y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] <= 0.05 and losses[-1] < 0.2 * losses[0], (losses[0], losses[-1])


def test_predictions_stay_same_across_calls_and_layout_cache():
    """Same batch, eval mode, two calls (second one reuses the cached device layout): identical outputs."""
    from chemprop_b200.nn import BondMessagePassing, MeanAggregation

    torch.manual_seed(1)
    bmg, _ = _data(32, seed=3)
    mp = BondMessagePassing(precision="bf16").cuda().eval()
    with torch.no_grad():
        a = MeanAggregation()(mp(bmg), bmg.batch)
        b = MeanAggregation()(mp(bmg), bmg.batch)
    assert torch.equal(a, b)            # deterministic kernels: bitwise reproducible
