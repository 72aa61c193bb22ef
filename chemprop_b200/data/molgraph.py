"""`MolGraph`: same fields and meaning as chemprop/data/molgraph.py:6-16."""
from typing import NamedTuple

import numpy as np


class MolGraph(NamedTuple):
    V: np.ndarray
    """``V x d_v`` atom features"""
    E: np.ndarray
    """``E x d_e`` bond features (one row per *directed* edge)"""
    edge_index: np.ndarray
    """``2 x E`` COO edges; row 0 = source atom, row 1 = destination atom"""
    rev_edge_index: np.ndarray
    """``E``; index of the reverse edge of each edge"""
