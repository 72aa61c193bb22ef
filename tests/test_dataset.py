"""CPU tier: the packed data set.  `PackedMolGraphDataset.batch(ids)` must hand back exactly the
BatchMolGraph the collate (chemprop/data/collate.py:37-62; pinned to the reference by tests/test_oracle.py's collate
fixture) builds from `[mgs[i] for i in ids]` -- every public tensor bit-identical, plus the compact transfer copy and
the host-computed layout meta words."""
import numpy as np
import pytest
import torch

from chemprop_b200 import DmpnnError
from chemprop_b200.data import BatchMolGraph, HostBatchBuffer, MolGraph, PackedMolGraphDataset, make_molecules
from oracle import restatement as R


def _same(a: BatchMolGraph, b: BatchMolGraph):
    assert len(a) == len(b)
    for k in ("V", "E", "edge_index", "rev_edge_index", "batch"):
        x, y = getattr(a, k), getattr(b, k)
        assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), k
    assert a._meta_host == b._meta_host
    assert (a._xfer is None) == (b._xfer is None)
    if a._xfer is not None:
        for x, y in zip(a._xfer, b._xfer):
            assert x.dtype == y.dtype and torch.equal(x.view(torch.int16) if x.dtype == torch.bfloat16 else x,
                                                      y.view(torch.int16) if y.dtype == torch.bfloat16 else y)


@pytest.mark.parametrize("kw", [dict(), dict(shuffle_edges=True, min_atoms=1), dict(mean_atoms=60, std_atoms=25, max_atoms=140)])
def test_batch_equals_collate_of_the_selected_molecules(kw):
    mgs = make_molecules(400, seed=21, **kw)
    ds = PackedMolGraphDataset.from_molgraphs(mgs)
    assert len(ds) == 400 and ds.d_v == 72 and ds.d_e == 14
    rng = np.random.default_rng(0)
    for ids in (np.arange(400), rng.permutation(400)[:150], rng.integers(0, 400, size=64), np.array([7]), np.array([], np.int64),
                np.array([3, 3, 3])):
        for compact in (None, torch.bfloat16):
            got = ds.batch(ids, transfer_dtype=compact)
            if len(ids) == 0:      # an empty selection keeps the feature widths (an empty collate cannot know them)
                assert len(got) == 0 and got.V.shape == (0, 72) and got.E.shape == (0, 14) and got.edge_index.shape == (2, 0)
                continue
            ref = BatchMolGraph([mgs[i] for i in ids], transfer_dtype=compact)
            _same(got, ref)
            # and against the oracle's restatement of the reference collate
            if len(ids):
                V, E, ei, rev, bt = R.collate([mgs[i] for i in ids])
                assert np.array_equal(got.V.numpy(), V) and np.array_equal(got.edge_index.numpy(), ei)
                assert np.array_equal(got.rev_edge_index.numpy(), rev) and np.array_equal(got.batch.numpy(), bt)


def test_meta_words_of_large_batches_match_the_numpy_layout():
    from chemprop_b200 import _lib
    from oracle import layout_np

    mgs = make_molecules(2500, seed=5, shuffle_edges=True, min_atoms=1)
    ds = PackedMolGraphDataset.from_molgraphs(mgs)
    ids = np.random.default_rng(1).permutation(2500)[:2100]
    b = ds.batch(ids)
    L = layout_np.build_layout(b.edge_index.numpy(), b.rev_edge_index.numpy(), b.batch.numpy(), len(b))
    m = b._meta_host
    assert (m[_lib.META_N_TILES], m[_lib.META_FLAGS], m[_lib.META_MAX_INDEG], m[_lib.META_MAX_TILE_ROWS],
            m[_lib.META_MAX_TILE_ATOMS]) == (L["n_tiles"], L["flags"], L["max_indeg"], L["max_tile_rows"], L["max_tile_atoms"])


def test_molgraph_round_trip_and_errors():
    mgs = make_molecules(30, seed=2, shuffle_edges=True)
    ds = PackedMolGraphDataset.from_molgraphs(mgs)
    for i in (0, 11, 29):
        mg = ds.molgraph(i)
        assert np.array_equal(mg.V, mgs[i].V) and np.array_equal(mg.E, mgs[i].E)
        assert np.array_equal(mg.edge_index, mgs[i].edge_index) and np.array_equal(mg.rev_edge_index, mgs[i].rev_edge_index)
    assert np.array_equal(ds.n_atoms([0, 5]), [mgs[0].V.shape[0], mgs[5].V.shape[0]])
    with pytest.raises(DmpnnError, match="out of range"):
        ds.batch([0, 30])
    with pytest.raises(DmpnnError, match="out of range"):
        ds.batch([-1])
    with pytest.raises(ValueError):
        ds.batch(np.zeros((2, 2), np.int64))
    m = mgs[0]
    broken = MolGraph(m.V, m.E, m.edge_index, np.roll(m.rev_edge_index, 1))
    with pytest.raises(DmpnnError, match="reverse-edge"):
        PackedMolGraphDataset.from_molgraphs([mgs[1], broken])


def test_pinned_batches_and_no_edge_molecules():
    rng = np.random.default_rng(3)
    from chemprop_b200.data import make_molecule

    mgs = [make_molecule(rng, n) for n in (1, 1, 5, 1, 9)]
    ds = PackedMolGraphDataset.from_molgraphs(mgs)
    _same(ds.batch([1, 0, 3]), BatchMolGraph([mgs[1], mgs[0], mgs[3]]))           # E = 0 batch
    _same(ds.batch([4, 1, 2]), BatchMolGraph([mgs[4], mgs[1], mgs[2]]))
    assert ds.nbytes() > 0


def test_reused_staging_buffer_gives_the_same_batches():
    mgs = make_molecules(300, seed=9, shuffle_edges=True, min_atoms=1)
    ds = PackedMolGraphDataset.from_molgraphs(mgs)
    rng = np.random.default_rng(4)
    for compact in (False, True):
        buf = HostBatchBuffer(ds.d_v, ds.d_e, atoms=100, edges=100, compact=compact)     # too small on purpose: it grows
        for n in (20, 250, 3, 120):
            ids = rng.integers(0, 300, size=n)
            got = ds.batch(ids, buffer=buf)
            ref = BatchMolGraph([mgs[i] for i in ids], transfer_dtype=torch.bfloat16 if compact else None)
            _same(got, ref)
            assert got.V.data_ptr() == buf.V.data_ptr() and got.edge_index.is_contiguous()
    with pytest.raises(ValueError):
        ds.batch([0], buffer=HostBatchBuffer(10, 3))
