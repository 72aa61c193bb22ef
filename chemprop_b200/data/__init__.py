from .collate import BatchMolGraph, Datum, TrainingBatch, collate_batch
from .dataset import HostBatchBuffer, PackedMolGraphDataset
from .molgraph import MolGraph
from .synthetic import make_chain_graph, make_cgr_graphs, make_molecule, make_molecules

__all__ = ["BatchMolGraph", "Datum", "TrainingBatch", "collate_batch", "MolGraph", "PackedMolGraphDataset", "HostBatchBuffer",
           "make_chain_graph", "make_cgr_graphs", "make_molecule", "make_molecules"]
