"""`MeanAggregation` / `SumAggregation` / `NormAggregation` with the API of chemprop/nn/agg.py:19-113
(`forward(H, batch) -> b x d`, `hparams`, registry names "mean" / "sum" / "norm"), executed as one
segmented reduction over the (contiguous) atom range of each molecule instead of
`index.repeat()` + `scatter_reduce_` (agg.py:74-78)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import _lib
from ..composed import GatherLinear
from ..engine import SegmentAggFunction, SegmentBcastFunction, segments_of

try:
    from .. import export as _export
except Exception:  # noqa: BLE001
    _export = None


class Aggregation(nn.Module):
    def __init__(self, dim: int = 0, *args, **kwargs):
        super().__init__()
        if dim != 0:
            raise NotImplementedError("the engine aggregates over dim 0 (atoms), as every chemprop model does")
        self.dim = dim
        self.hparams = {"dim": dim, "cls": self.__class__}

    _mode = _lib.SCALE_NONE
    _scale = 1.0

    def forward(self, H: Tensor, batch: Tensor) -> Tensor:
        if _export is not None and _export.is_tracing():      # torch.export: one custom op (chemprop_b200/export.py)
            return torch.ops.dmpnn.segment_agg(H, batch, int(self._mode), float(self._scale))
        ptr, seg_of_row, B = segments_of(batch)
        return SegmentAggFunction.apply(H, ptr, seg_of_row, B, self._mode, float(self._scale))


class MeanAggregation(Aggregation):
    _mode = _lib.SCALE_INV_COUNT


class SumAggregation(Aggregation):
    _mode = _lib.SCALE_NONE


class NormAggregation(SumAggregation):
    _mode = _lib.SCALE_DIV_CONST

    def __init__(self, dim: int = 0, *args, norm: float = 100.0, **kwargs):
        super().__init__(dim, **kwargs)
        self.norm = norm
        self._scale = norm
        self.hparams["norm"] = norm


class AttentiveAggregation(Aggregation):
    """chemprop/nn/agg.py:116-133: `alpha_v = exp(W h_v) / sum_{u in mol(v)} exp(W h_u)`, `h = sum_v alpha_v h_v` -- a
    segmented softmax-weighted sum.  Composed from the engine's segment primitives: the `d -> 1` logit
    layer (dmpnn_linear_fwd), the per-molecule sum of the exponentials (dmpnn_segment_sum), its broadcast back to the atoms
    (dmpnn_segment_bcast) and the weighted per-molecule sum; exp / divide / multiply are torch elementwise ops in
    between, exactly the reference's op sequence (no max-subtraction there either)."""

    def __init__(self, dim: int = 0, *args, output_size: int, **kwargs):
        super().__init__(dim, *args, **kwargs)
        self.hparams["output_size"] = output_size
        self.W = nn.Linear(output_size, 1)

    def forward(self, H: Tensor, batch: Tensor) -> Tensor:
        ptr, seg_of_row, B = segments_of(batch)
        Hf = H.float()
        logits = GatherLinear.apply(Hf, None, None, None, self.W.weight, self.W.bias, None, Hf.shape[0]).exp()   # agg.py:123
        Z = SegmentAggFunction.apply(logits, ptr, seg_of_row, B, _lib.SCALE_NONE, 1.0)                            # agg.py:124-126
        alphas = logits / SegmentBcastFunction.apply(Z, ptr, seg_of_row, Hf.shape[0])                             # agg.py:127
        return SegmentAggFunction.apply(alphas * Hf, ptr, seg_of_row, B, _lib.SCALE_NONE, 1.0)                    # agg.py:128-131


AggregationRegistry = {"mean": MeanAggregation, "sum": SumAggregation, "norm": NormAggregation}
