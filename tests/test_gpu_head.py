"""GPU tier: the molecule-level head on the engine's kernels (csrc/head.cu: batch norm, MSE criterion; FFN GEMMs with the
activation in the epilogue) and the whole training step as ONE CUDA graph (chemprop_b200/graph.py) -- SURVEY.md section 8f-2,
against the reference's own `MPNN.training_step` (tests/golden/fixture_mpnn_head.npz, oracle/make_golden.py::mpnn_head_case)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_engine_mpnn_head_matches_reference_training_step():
    from tests.util import check_mpnn_head

    check_mpnn_head("cuda")


def test_training_step_as_one_cuda_graph_matches_reference():
    from tests.util import check_mpnn_head

    check_mpnn_head("cuda", graph=True)


@pytest.mark.parametrize("B,d", [(1, 40), (7, 300), (10000, 300), (513, 33)])
def test_batch_norm_kernels_vs_torch(B, d):
    from chemprop_b200.nn import EngineBatchNorm1d

    torch.manual_seed(B + d)
    ours, ref = EngineBatchNorm1d(d).cuda(), torch.nn.BatchNorm1d(d).cuda()
    with torch.no_grad():
        ours.weight.uniform_(0.5, 1.5); ours.bias.normal_()
        ref.load_state_dict(ours.state_dict())
    x = (torch.randn(B, d, device="cuda") * 3 + 1).requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    G = torch.randn(B, d, device="cuda")
    if B == 1:
        with pytest.raises(ValueError):
            ref(x2)                                  # torch refuses a training batch of one; the engine normalises it (var = 0)
        return
    (ours(x) * G).sum().backward()
    (ref(x2) * G).sum().backward()
    torch.testing.assert_close(x.grad, x2.grad, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(ours.weight.grad, ref.weight.grad, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(ours.bias.grad, ref.bias.grad, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(ours.running_mean, ref.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ours.running_var, ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(ours.num_batches_tracked) == 1


def test_cuda_graph_step_replays_new_batches_of_the_same_signature():
    """Graph replay with NEW data: same molecules' topology (same signature), different features and targets -> equals eager."""
    from chemprop_b200.data import BatchMolGraph, make_molecules
    from chemprop_b200.graph import CudaGraphStep
    from chemprop_b200.nn import BondMessagePassing, EngineMPNN

    torch.manual_seed(0)
    mgs = make_molecules(600, seed=3)
    model = EngineMPNN(BondMessagePassing(precision="bf16"), batch_norm=True).cuda().train()
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    Y = torch.randn(600, 1, device="cuda")

    def fn(b):
        for p in model.parameters():
            p.grad.zero_()
        loss = model.training_loss(b, Y)
        loss.backward()
        return loss

    step = CudaGraphStep(fn)
    a = BatchMolGraph(mgs)
    a.to("cuda")
    step(a)
    b = BatchMolGraph(mgs)
    b.V = b.V * 0.5 + 0.1                           # new features, same topology
    b.to("cuda")
    Y.copy_(torch.randn(600, 1, device="cuda"))
    bn_state = {k: v.clone() for k, v in model.bn.state_dict().items()}
    loss_g = step(b).clone()
    grads_g = {k: p.grad.clone() for k, p in model.named_parameters()}
    assert step.captures == 1 and step.replays == 2
    model.bn.load_state_dict(bn_state)
    loss_e = fn(b)
    torch.testing.assert_close(loss_g, loss_e, rtol=1e-5, atol=1e-6)
    for k, p in model.named_parameters():
        torch.testing.assert_close(grads_g[k], p.grad, rtol=1e-4, atol=1e-6, msg=k)
    c = BatchMolGraph(make_molecules(500, seed=4))   # another signature: captured on first sight
    c.to("cuda")
    Y2 = Y[:500].clone()
    model2_loss = None
    step2 = CudaGraphStep(lambda bb: fn_sized(bb))

    def fn_sized(bb):
        for p in model.parameters():
            p.grad.zero_()
        loss = model.training_loss(bb, Y2 if len(bb) == 500 else Y)
        loss.backward()
        return loss

    l1 = step2(c).clone()
    l2 = step2(b).clone()
    assert step2.captures == 2 and torch.isfinite(l1) and torch.isfinite(l2)
