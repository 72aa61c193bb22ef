"""CPU tier: torch.export of a model built on the engine's modules, mirroring the reference's
tests/integration/test_export.py:15-46 (dynamic atom / edge counts; run on a batch WITHOUT edges).  Tracing needs no GPU
(the custom ops' fake implementations carry the shapes); the exported program is then executed with the kernel
wrappers emulated (tests/emu.py) and must equal the eager result and the oracle.  Without the emulation the exported
program refuses CPU tensors like everything else (no fallback)."""
import numpy as np
import pytest
import torch

from chemprop_b200 import DmpnnError
from chemprop_b200.data import BatchMolGraph, make_molecule, make_molecules
from chemprop_b200.export import register_batch_mol_graph_pytree
from chemprop_b200.nn import AtomMessagePassing, BondMessagePassing, MeanAggregation, SumAggregation
from oracle import restatement as R
from tests import emu


class _Model(torch.nn.Module):
    def __init__(self, mp, agg):
        super().__init__()
        self.message_passing, self.agg = mp, agg
        self.head = torch.nn.Linear(mp.output_dim, 1)

    def forward(self, bmg):
        return self.head(self.agg(self.message_passing(bmg), bmg.batch).float())


@pytest.mark.parametrize("kind,agg_cls,mode", [("bond", SumAggregation, "sum"), ("atom", MeanAggregation, "mean")])
def test_export_with_dynamic_graph_sizes(kind, agg_cls, mode, monkeypatch):
    register_batch_mol_graph_pytree()
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    export_graph = BatchMolGraph(make_molecules(4, seed=1, mean_atoms=6, std_atoms=2))
    inference_graph = BatchMolGraph([make_molecule(rng, 1) for _ in range(4)])            # "C", "S", "N", "O": no edges
    other_graph = BatchMolGraph(make_molecules(4, seed=2, mean_atoms=9, std_atoms=3, shuffle_edges=True))
    assert export_graph.V.shape[0] != inference_graph.V.shape[0] and inference_graph.E.shape[0] == 0
    cls = BondMessagePassing if kind == "bond" else AtomMessagePassing
    model = _Model(cls(d_h=24, depth=3, activation="tanh", bias=True), agg_cls()).eval()
    num_atoms, num_edges = torch.export.Dim("num_atoms"), torch.export.Dim("num_edges")
    dynamic_shapes = {"bmg": [{0: num_atoms}, {0: num_edges}, {1: num_edges}, {0: num_edges}, {0: num_atoms}]}
    exported = torch.export.export(model, (export_graph,), dynamic_shapes=dynamic_shapes, strict=False)   # no GPU needed
    targets = {str(n.target) for n in exported.graph.nodes if n.op == "call_function"}
    assert "dmpnn.mp_forward.default" in targets and "dmpnn.segment_agg.default" in targets
    with pytest.raises((DmpnnError, RuntimeError), match="no CPU fallback|CUDA"):
        exported.module()(inference_graph)                                             # the real op: CUDA only
    emu.patch_engine(monkeypatch)
    P = {k: v.detach() for k, v in model.message_passing.state_dict().items()}
    for g in (inference_graph, other_graph, export_graph):
        with torch.inference_mode():
            expected = model(g)
            actual = exported.module()(g)
        torch.testing.assert_close(actual, expected)
        H = R.message_passing_forward(kind, g.V, g.E, g.edge_index, g.rev_edge_index, P["W_i.weight"], P["W_i.bias"],
                                      P["W_h.weight"], P["W_h.bias"], P["W_o.weight"], P["W_o.bias"], 3, "tanh")
        ref = model.head(R.aggregate(H, g.batch, mode, n_mols=4))
        torch.testing.assert_close(actual, ref, rtol=1e-5, atol=1e-5)


def test_export_refuses_the_composed_tier():
    register_batch_mol_graph_pytree()
    model = _Model(BondMessagePassing(d_h=8, activation="prelu"), SumAggregation()).eval()
    with pytest.raises(Exception, match="monolithic tiers"):
        torch.export.export(model, (BatchMolGraph(make_molecules(2, seed=0)),), strict=False)
