from .collate import BatchMolGraph, Datum, TrainingBatch, collate_batch
from .molgraph import MolGraph
from .synthetic import make_chain_graph, make_cgr_graphs, make_molecule, make_molecules

__all__ = ["BatchMolGraph", "Datum", "TrainingBatch", "collate_batch", "MolGraph",
           "make_chain_graph", "make_cgr_graphs", "make_molecule", "make_molecules"]
