"""TEST INFRASTRUCTURE ONLY.  SMILES -> molecular TOPOLOGY for BASELINE config 1 (the reference's own CPU case:
tests/data/regression.csv, batch 50).  RDKit is absent here and on the GPU box, so the reference's featuriser
(chemprop/featurizers/molgraph/molecule.py:75-90) cannot run; this ~100-line parser covers the subset of SMILES that file
uses (organic-subset atoms, aromatic lower case, `[nH]`-style brackets, branches, ring-closure digits, bond symbols
- = # : / \\) and produces MolGraphs in the featuriser's EDGE ORDER (bond i -> directed edges 2i: u->v, 2i+1: v->u, the two
sharing one feature row; rev = e ^ 1; molecule.py:81-89) with stand-in multi-hot features of the reference's widths
(72 atom / 14 bond).  Parity on the message-passing path does not depend on the chemical correctness of the FEATURES:
the reference and the engine consume the same BatchMolGraph (hard part 7).
"""
from __future__ import annotations

import numpy as np

ELEMENTS = ["C", "N", "O", "S", "F", "Cl", "Br", "I", "P", "B", "Si", "Se", "H"]
MASS = {"C": 12.011, "N": 14.007, "O": 15.999, "S": 32.06, "F": 18.998, "Cl": 35.45, "Br": 79.904, "I": 126.904, "P": 30.974,
        "B": 10.81, "Si": 28.085, "Se": 78.971, "H": 1.008}
BOND_ORDER = {"-": 0, "=": 1, "#": 2, ":": 3, "/": 0, "\\": 0}


def parse(smiles: str):
    """-> (atoms: list of (element, aromatic, n_explicit_H), bonds: list of (u, v, order_index)) in order of appearance
    (ring-closure bonds appear when the ring CLOSES, as in RDKit's atom / bond numbering of a parsed SMILES)."""
    atoms, bonds = [], []
    stack, prev, pending = [], None, None
    rings: dict[str, tuple[int, int | None]] = {}
    i, n = 0, len(smiles)

    def add_atom(sym: str, aromatic: bool, nH: int = 0):
        nonlocal prev, pending
        atoms.append((sym, aromatic, nH))
        a = len(atoms) - 1
        if prev is not None:
            order = pending if pending is not None else (3 if (aromatic and atoms[prev][1]) else 0)
            bonds.append((prev, a, order))
        prev, pending = a, None

    while i < n:
        c = smiles[i]
        if c == "[":
            j = smiles.index("]", i)
            body = smiles[i + 1:j].lstrip("0123456789")
            sym = body[:2] if body[:2] in ELEMENTS else body[:1]
            aromatic = sym.islower()
            rest = body[len(sym):]
            nH = 0
            if "H" in rest:
                k = rest.index("H") + 1
                nH = int(rest[k]) if k < len(rest) and rest[k].isdigit() else 1
            add_atom(sym.capitalize(), aromatic, nH)
            i = j + 1
        elif c in "BCNOPSFI" or c in "bcnops":
            two = smiles[i:i + 2]
            if two in ("Cl", "Br"):
                add_atom(two, False)
                i += 2
            else:
                add_atom(c.upper(), c.islower())
                i += 1
        elif c in BOND_ORDER:
            pending = BOND_ORDER[c]
            i += 1
        elif c == "(":
            stack.append(prev)
            i += 1
        elif c == ")":
            prev = stack.pop()
            i += 1
        elif c.isdigit() or c == "%":
            key = smiles[i + 1:i + 3] if c == "%" else c
            i += 3 if c == "%" else 1
            if key in rings:
                a, order0 = rings.pop(key)
                order = pending if pending is not None else order0
                if order is None:
                    order = 3 if (atoms[a][1] and atoms[prev][1]) else 0
                bonds.append((a, prev, order))
            else:
                rings[key] = (prev, pending)
            pending = None
        elif c == ".":
            prev, pending = None, None
            i += 1
        else:
            raise ValueError(f"unsupported SMILES character {c!r} in {smiles!r}")
    if rings or stack:
        raise ValueError(f"unbalanced ring closures / branches in {smiles!r}")
    return atoms, bonds


def to_molgraph(smiles: str, d_v: int = 72, d_e: int = 14):
    """Stand-in featurisation with the reference's widths: atom = one-hot element (13) + one-hot degree (6) + aromatic +
    explicit-H count + mass * 0.01 in the last column; bond = one-hot order (4) + in-ring flag + both-aromatic flag."""
    from chemprop_b200.data import MolGraph

    atoms, bonds = parse(smiles)
    n, nb = len(atoms), len(bonds)
    deg = np.zeros(n, dtype=np.int64)
    for u, v, _ in bonds:
        deg[u] += 1
        deg[v] += 1
    V = np.zeros((n, d_v), dtype=np.float32)
    for a, (sym, aromatic, nH) in enumerate(atoms):
        V[a, ELEMENTS.index(sym)] = 1
        V[a, 13 + min(int(deg[a]), 5)] = 1
        V[a, 19] = float(aromatic)
        V[a, 20 + min(nH, 3)] = 1
        V[a, d_v - 1] = 0.01 * MASS[sym]
    # a bond is in a ring iff removing it keeps its endpoints connected: n is tiny, a DFS per bond is fine
    adj = [[] for _ in range(n)]
    for b, (u, v, _) in enumerate(bonds):
        adj[u].append((v, b))
        adj[v].append((u, b))

    def connected_without(b, u, v):
        seen, todo = {u}, [u]
        while todo:
            x = todo.pop()
            for y, bb in adj[x]:
                if bb != b and y not in seen:
                    seen.add(y)
                    todo.append(y)
        return v in seen

    E = np.zeros((2 * nb, d_e), dtype=np.float32)
    edge_index = np.zeros((2, 2 * nb), dtype=np.int64)
    for b, (u, v, order) in enumerate(bonds):
        f = np.zeros(d_e, dtype=np.float32)
        f[order] = 1
        f[4] = float(connected_without(b, u, v))
        f[5] = float(atoms[u][1] and atoms[v][1])
        E[2 * b] = E[2 * b + 1] = f                                     # molecule.py:84-86: both directions share the row
        edge_index[:, 2 * b] = (u, v)
        edge_index[:, 2 * b + 1] = (v, u)
    rev = np.arange(2 * nb).reshape(-1, 2)[:, ::-1].ravel()            # molecule.py:89
    return MolGraph(V, E, edge_index, rev.astype(np.int64))
