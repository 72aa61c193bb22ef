from .agg import Aggregation, AggregationRegistry, MeanAggregation, NormAggregation, SumAggregation
from .message_passing import AtomMessagePassing, BondMessagePassing
from .transforms import GraphTransform, ScaleTransform

__all__ = ["Aggregation", "AggregationRegistry", "MeanAggregation", "NormAggregation", "SumAggregation",
           "AtomMessagePassing", "BondMessagePassing", "GraphTransform", "ScaleTransform"]
