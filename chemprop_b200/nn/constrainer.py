"""`ConstrainerFFN` with the API of chemprop/nn/ffn.py:70-145: atom- or bond-level predictions are adjusted
so that they sum to a per-molecule constraint, the adjustment being shared out with a per-molecule softmax of an MLP's
output.  The MLP is the reference's plain torch stack (same module tree, so state-dict keys `ffn.<i>.<j>.*` match); the
segmented part -- per-molecule sum of exp(k), its broadcast back to the atoms / bonds, the per-molecule sum of the
predictions, the broadcast of the deviation -- runs on the engine's segment primitives (`dmpnn_segment_sum`,
`dmpnn_segment_bcast`), with exp / divide / multiply as torch elementwise ops in between (the reference's op sequence)."""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor, nn

from .. import _lib
from ..engine import SegmentAggFunction, SegmentBcastFunction, segments_of
from .message_passing import DEFAULT_HIDDEN_DIM, get_activation_function


def build_mlp(input_dim: int, output_dim: int, hidden_dim: int | Sequence[int] = 300, n_layers: int = 1,
              dropout: float = 0.0, activation="relu") -> nn.Sequential:
    """Module tree of chemprop/nn/ffn.py `MLP.build` (:38-61): [Linear], then [act, dropout, Linear] per further layer."""
    drop, act = nn.Dropout(dropout), get_activation_function(activation)
    hidden = [hidden_dim] * n_layers if isinstance(hidden_dim, int) else list(hidden_dim)
    dims = [input_dim] + hidden + [output_dim]
    blocks = [nn.Sequential(nn.Linear(dims[0], dims[1]))]
    blocks += [nn.Sequential(act, drop, nn.Linear(d1, d2)) for d1, d2 in zip(dims[1:-1], dims[2:])]
    return nn.Sequential(*blocks)


class ConstrainerFFN(nn.Module):
    def __init__(self, n_constraints: int = 1, fp_dim: int = DEFAULT_HIDDEN_DIM, hidden_dim: int | Sequence[int] = 300,
                 n_layers: int = 1, dropout: float = 0.0, activation="relu"):
        super().__init__()
        self.hparams = dict(n_constraints=n_constraints, fp_dim=fp_dim, hidden_dim=hidden_dim, n_layers=n_layers,
                            dropout=dropout, activation=activation, cls=self.__class__)
        self.ffn = build_mlp(fp_dim, n_constraints, hidden_dim, n_layers, dropout, activation)

    def forward(self, fp: Tensor, preds: Tensor, batch: Tensor, constraints: Tensor) -> Tensor:
        """fp: b x h fingerprints; preds: b x t; batch: b (molecule of each atom / bond, sorted); constraints: m x t (NaN in
        row 0 marks an unconstrained column, ffn.py:137).  Returns the adjusted b x t predictions."""
        # ffn.py:123 sizes everything by the constraints' row count: molecules without an atom / bond row (a trailing
        # single-heavy-atom SMILES has no bond) are empty segments, not an error
        n_mols = int(constraints.shape[0])
        ptr, seg, _ = segments_of(batch, n_seg=n_mols)
        rows = preds.shape[0]
        has = ~torch.isnan(constraints)[0]                                                           # ffn.py:136
        if rows == 0 or not bool(has.any()):             # nothing to adjust (the segment kernels need >= 1 column)
            return preds + torch.zeros_like(preds)
        expk = self.ffn(fp).exp()                                                                  # ffn.py:120-121
        Z = SegmentAggFunction.apply(expk.float(), ptr, seg, n_mols, _lib.SCALE_NONE, 1.0)           # ffn.py:124-127
        w = expk / SegmentBcastFunction.apply(Z, ptr, seg, rows)                                     # ffn.py:128-129
        per_mol = SegmentAggFunction.apply(preds.float(), ptr, seg, n_mols, _lib.SCALE_NONE, 1.0)    # ffn.py:131-134
        deviation = (constraints[:, has] - per_mol[:, has]).contiguous()                             # ffn.py:137
        corrections = w * SegmentBcastFunction.apply(deviation, ptr, seg, rows)                      # ffn.py:139
        out = torch.zeros_like(preds)
        out[:, has] = corrections.to(preds.dtype)                                                    # ffn.py:140-141
        return preds + out
