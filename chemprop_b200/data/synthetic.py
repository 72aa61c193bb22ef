"""Seeded synthetic molecule generator: drug-like random trees plus ring
closures, emitted in the reference featuriser's edge order (edges 2i: u->v, 2i+1: v->u,
rev[e] = e ^ 1; chemprop/featurizers/molgraph/molecule.py:75-90)."""
from __future__ import annotations

import numpy as np

from .molgraph import MolGraph


def make_molecule(rng: np.random.Generator, n_atoms: int, d_v: int = 72, d_e: int = 14,
                  ring_frac: float = 0.06, max_deg: int = 4, shuffle_edges: bool = False) -> MolGraph:
    n = int(n_atoms)
    deg = np.zeros(n, dtype=np.int64)
    bonds = []
    for v in range(1, n):
        lo = max(0, v - 6)
        cand = [u for u in range(lo, v) if deg[u] < max_deg]
        if not cand:
            cand = [u for u in range(0, v) if deg[u] < max_deg] or [v - 1]
        u = int(cand[rng.integers(len(cand))])
        bonds.append((u, v))
        deg[u] += 1
        deg[v] += 1
    n_ring = int(rng.binomial(n, ring_frac)) if n > 3 else 0
    have = set(bonds)
    for _ in range(n_ring):
        u, v = (int(x) for x in rng.integers(0, n, size=2))
        if u == v:
            continue
        a, b = min(u, v), max(u, v)
        if (a, b) in have or deg[a] >= max_deg or deg[b] >= max_deg:
            continue
        bonds.append((a, b))
        have.add((a, b))
        deg[a] += 1
        deg[b] += 1
    nb = len(bonds)
    V = np.zeros((n, d_v), dtype=np.float32)
    if d_v > 1:
        cols = rng.integers(0, d_v - 1, size=(n, 7))
        V[np.arange(n)[:, None], cols] = 1.0
    V[:, -1] = rng.uniform(0.01, 0.4, size=n).astype(np.float32)
    Eb = np.zeros((nb, d_e), dtype=np.float32)
    if nb:
        cols = rng.integers(0, d_e, size=(nb, 3))
        Eb[np.arange(nb)[:, None], cols] = 1.0
    E = np.repeat(Eb, 2, axis=0)
    ei = np.zeros((2, 2 * nb), dtype=np.int64)
    for i, (u, v) in enumerate(bonds):
        ei[:, 2 * i] = (u, v)
        ei[:, 2 * i + 1] = (v, u)
    rev = np.arange(2 * nb, dtype=np.int64).reshape(-1, 2)[:, ::-1].ravel() if nb else np.zeros(0, dtype=np.int64)
    if shuffle_edges and nb:
        p = rng.permutation(2 * nb)          # new position j holds old edge p[j]
        inv = np.empty_like(p)
        inv[p] = np.arange(2 * nb)
        ei, E, rev = ei[:, p], E[p], inv[rev[p]]
    return MolGraph(V=V, E=E, edge_index=ei, rev_edge_index=rev)


def make_molecules(n_mols: int, seed: int = 0, mean_atoms: float = 25.0, std_atoms: float = 5.0,
                   min_atoms: int = 2, max_atoms: int = 60, d_v: int = 72, d_e: int = 14,
                   ring_frac: float = 0.06, shuffle_edges: bool = False) -> list[MolGraph]:
    rng = np.random.default_rng(seed)
    sizes = np.clip(np.rint(rng.normal(mean_atoms, std_atoms, size=n_mols)), min_atoms, max_atoms).astype(int)
    return [make_molecule(rng, int(s), d_v, d_e, ring_frac, shuffle_edges=shuffle_edges) for s in sizes]


def make_cgr_graphs(n_graphs: int, seed: int = 0, d_v: int = 106, d_e: int = 28) -> list[MolGraph]:
    """~80-atom near-tree condensed reaction graphs (BASELINE config 4)."""
    return make_molecules(n_graphs, seed, mean_atoms=80.0, std_atoms=10.0, min_atoms=20, max_atoms=120,
                          d_v=d_v, d_e=d_e, ring_frac=0.03)


def make_chain_graph(num_atoms: int, d_v: int = 72, d_e: int = 14) -> MolGraph:
    """The block-ordered chain of the reference's unit test (tests/unit/nn/test_message_passing.py:15-26):
    all forward edges first, then all backward edges (so rev[e] != e ^ 1)."""
    src = np.arange(num_atoms - 1)
    dst = src + 1
    edge_index = np.stack((np.concatenate((src, dst)), np.concatenate((dst, src))))
    num_edges = edge_index.shape[1]
    return MolGraph(
        V=np.ones((num_atoms, d_v), dtype=np.float32),
        E=np.ones((num_edges, d_e), dtype=np.float32),
        edge_index=edge_index,
        rev_edge_index=np.concatenate((np.arange(num_atoms - 1, num_edges), src)),
    )
