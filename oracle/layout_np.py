"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the engine's integer layout
(dmpnn_layout_build): stable sort of edges by destination atom, per-row src / dst / rev, molecule
offsets and the greedy molecule-aligned tile table.  Integer work -> the CUDA output must be
bit-exact with this."""
from __future__ import annotations

import numpy as np

TILE_ROWS = 128
TILE_ATOMS = 128
TILE_CHUNK = 1024   # greedy packing restarts every TILE_CHUNK molecules (chunks are packed independently)


def build_layout(edge_index: np.ndarray, rev: np.ndarray, batch: np.ndarray, n_mols: int) -> dict:
    src, dst = edge_index[0].astype(np.int64), edge_index[1].astype(np.int64)
    E, V, B = src.shape[0], batch.shape[0], int(n_mols)
    perm = np.argsort(dst, kind="stable").astype(np.int32)
    inv_perm = np.empty(E, dtype=np.int32)
    inv_perm[perm] = np.arange(E, dtype=np.int32)
    rowptr = np.zeros(V + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=V), out=rowptr[1:])
    src_row = src[perm].astype(np.int32)
    dst_row = dst[perm].astype(np.int32)
    rev_row = inv_perm[rev[perm]].astype(np.int32) if E else np.zeros(0, np.int32)
    mol_atom_ptr = np.searchsorted(batch, np.arange(B + 1), side="left").astype(np.int32)
    mol_row_ptr = rowptr[mol_atom_ptr].astype(np.int32)
    tiles = []
    t_mol, max_rows, max_atoms = 0, 0, 0
    for m in range(B):
        if m > t_mol and (m % TILE_CHUNK == 0 or mol_row_ptr[m + 1] - mol_row_ptr[t_mol] > TILE_ROWS
                          or mol_atom_ptr[m + 1] - mol_atom_ptr[t_mol] > TILE_ATOMS):
            tiles.append(t_mol)
            max_rows = max(max_rows, int(mol_row_ptr[m] - mol_row_ptr[t_mol]))
            max_atoms = max(max_atoms, int(mol_atom_ptr[m] - mol_atom_ptr[t_mol]))
            t_mol = m
    if B > 0:
        tiles.append(t_mol)
        max_rows = max(max_rows, int(mol_row_ptr[B] - mol_row_ptr[t_mol]))
        max_atoms = max(max_atoms, int(mol_atom_ptr[B] - mol_atom_ptr[t_mol]))
    n_tiles = len(tiles)
    tile_mol_ptr = np.array(tiles + [B], dtype=np.int32)
    tile_row_ptr = mol_row_ptr[tile_mol_ptr].astype(np.int32)
    tile_atom_ptr = mol_atom_ptr[tile_mol_ptr].astype(np.int32)
    # validity flags
    in_range = bool(E == 0 or (src.min() >= 0 and src.max() < V and dst.min() >= 0 and dst.max() < V
                               and rev.min() >= 0 and rev.max() < E))
    in_range = in_range and bool(V == 0 or (batch.min() >= 0 and batch.max() < B))
    invol = in_range and bool(E == 0 or (np.array_equal(rev[rev], np.arange(E)) and np.array_equal(src[rev], dst)
                                         and np.array_equal(dst[rev], src)))
    sorted_ = in_range and bool(np.all(np.diff(batch) >= 0)) and bool(E == 0 or np.array_equal(batch[src], batch[dst]))
    flags = (1 if invol else 0) | (2 if sorted_ else 0) | (4 if in_range else 0)
    max_indeg = int(np.diff(rowptr).max()) if V else 0
    return dict(perm=perm, inv_perm=inv_perm, rowptr=rowptr, src_row=src_row, dst_row=dst_row, rev_row=rev_row,
                mol_atom_ptr=mol_atom_ptr, mol_row_ptr=mol_row_ptr, tile_mol_ptr=tile_mol_ptr, tile_row_ptr=tile_row_ptr,
                tile_atom_ptr=tile_atom_ptr, n_tiles=n_tiles,
                flags=flags, max_indeg=max_indeg, max_tile_rows=max_rows, max_tile_atoms=max_atoms)
