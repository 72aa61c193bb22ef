"""`MeanAggregation` / `SumAggregation` / `NormAggregation` with the API of chemprop/nn/agg.py:19-113
(`forward(H, batch) -> b x d`, `hparams`, registry names "mean" / "sum" / "norm"), executed as one
segmented reduction over the (contiguous) atom range of each molecule instead of
`index.repeat()` + `scatter_reduce_` (agg.py:74-78)."""
from __future__ import annotations

from torch import Tensor, nn

from .. import _lib
from ..engine import SegmentAggFunction, segments_of


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


AggregationRegistry = {"mean": MeanAggregation, "sum": SumAggregation, "norm": NormAggregation}
