"""CPU tier: the C-ABI library loads and exports every symbol include/dmpnn.h declares
(no compute is launched without a GPU), and the host-side entry point works."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from chemprop_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dmpnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmpnn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dmpnn.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_binding_matches_header_prototypes():
    """Every ctypes signature has the parameter count and the scalar kinds (pointer / 64-bit / 32-bit / float) of the
    prototype in include/dmpnn.h -- an ABI drift between header and binding would corrupt the call silently."""
    text = open(os.path.join(ROOT, "include", "dmpnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(dmpnn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert set(protos) == set(_lib.SIGNATURES)

    def kind_of_c(param: str) -> str:
        param = param.strip()
        if "*" in param:
            return "ptr"
        if re.search(r"\b(int64_t|uint64_t|size_t)\b", param):
            return "i64"
        if re.search(r"\bfloat\b", param):
            return "f32"
        if re.search(r"\bint\b", param):
            return "i32"
        raise AssertionError(f"unrecognised parameter {param!r}")

    def kind_of_ctypes(t) -> str:
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or hasattr(t, "_type_") and isinstance(t._type_, type):
            return "ptr"
        return {ctypes.c_int64: "i64", ctypes.c_uint64: "i64", ctypes.c_size_t: "i64", ctypes.c_int: "i32", ctypes.c_float: "f32"}[t]

    for name, params in protos.items():
        plist = [] if params.strip() in ("", "void") else [x for x in params.split(",")]
        _, argtypes = _lib.SIGNATURES[name]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes))
        for i, (cp, at) in enumerate(zip(plist, argtypes)):
            assert kind_of_c(cp) == kind_of_ctypes(at), (name, i, cp.strip(), at)


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.dmpnn_version() == 100
    assert isinstance(lib.dmpnn_last_error(), bytes)
    assert lib.dmpnn_device_ok() in (0, 1)


def test_bad_arguments_fail_loudly():
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    assert lib.dmpnn_layout_workspace_bytes(-1, 0, 0, ctypes.byref(n)) != 0
    assert b"bad args" in lib.dmpnn_last_error()
    with pytest.raises(_lib.DmpnnError):
        _lib.check(lib.dmpnn_linear_wgrad_workspace_bytes(1, 0, 1, ctypes.byref(n)), "wgrad_workspace_bytes")


def test_new_entry_points_validate_their_arguments_before_touching_the_device():
    """Argument checks of the device-side entry points added for the packed data set / dropout run on the host, before
    any CUDA call: they are callable (and must fail cleanly) without a GPU."""
    lib = _lib.load()
    assert lib.dmpnn_dataset_gather(None, None, None, -1, None, None, None, None, None, None, 0, 72, 14,
                                    None, None, None, None, None, 0, None) != 0
    assert b"dataset_gather" in lib.dmpnn_last_error()
    assert lib.dmpnn_dataset_gather(None, None, None, 0, None, None, None, None, None, None, 0, 72, 14,
                                    None, None, None, None, None, 0, None) == 0            # empty selection: nothing to do
    assert lib.dmpnn_dataset_gather(None, None, None, 3, None, None, None, None, None, None, 0, 72, 14,
                                    None, None, None, None, None, 0, None) != 0            # null tables
    assert lib.dmpnn_scale_mask(None, None, None, _lib.F32, 5, 1.0, None) != 0 and b"scale_mask" in lib.dmpnn_last_error()
    assert lib.dmpnn_scale_mask(None, None, None, 7, 0, 1.0, None) != 0                     # unknown dtype
    assert lib.dmpnn_scale_mask(None, None, None, _lib.BF16, 0, 1.0, None) == 0
    meta = np.zeros(_lib.META_WORDS, dtype=np.int32)
    assert lib.dmpnn_batch_meta_host(None, None, None, 0, 0, 0, meta.ctypes.data) == 0 and meta[_lib.META_FLAGS] == 7
    assert lib.dmpnn_batch_meta_host(None, None, None, 3, 0, 1, meta.ctypes.data) != 0      # atoms without a batch array
    ids = np.array([5], dtype=np.int64)
    ptr = np.array([0, 2], dtype=np.int64)
    out = np.zeros(2, dtype=np.int64)
    mi = np.zeros(1, dtype=np.int32)
    assert lib.dmpnn_dataset_batch_meta_host(1, ids.ctypes.data, 1, ptr.ctypes.data, ptr.ctypes.data, mi.ctypes.data,
                                             out.ctypes.data, out.ctypes.data, meta.ctypes.data) != 0
    assert b"out of range" in lib.dmpnn_last_error()
    assert lib.dmpnn_tile_pack_order(-1, None, None, None) != 0


def test_collate_host_matches_oracle_and_reference_fixture():
    from chemprop_b200.data import BatchMolGraph, MolGraph, make_molecules
    from oracle import restatement as R
    from tests.util import load_golden

    g = load_golden("collate_fixture")
    mgs = [MolGraph(g[f"mg{i}.V"], g[f"mg{i}.E"], g[f"mg{i}.edge_index"], g[f"mg{i}.rev_edge_index"]) for i in range(2)]
    bmg = BatchMolGraph(mgs)
    for k in ("V", "E", "edge_index", "rev_edge_index", "batch"):
        got = getattr(bmg, k)
        assert got.numpy().dtype == g[k].dtype and np.array_equal(got.numpy(), g[k]), k
    assert len(bmg) == 2
    # ragged / shuffled / single-atom molecules
    mgs = make_molecules(300, seed=3, shuffle_edges=True, min_atoms=1)
    bmg = BatchMolGraph(mgs)
    V, E, ei, rev, batch = R.collate(mgs)
    assert np.array_equal(bmg.V.numpy(), V) and np.array_equal(bmg.E.numpy(), E)
    assert np.array_equal(bmg.edge_index.numpy(), ei) and np.array_equal(bmg.rev_edge_index.numpy(), rev)
    assert np.array_equal(bmg.batch.numpy(), batch)
    assert bmg.V.dtype == torch.float32 and bmg.edge_index.dtype == torch.int64


def test_collate_empty_edges():
    from chemprop_b200.data import BatchMolGraph, make_molecule

    rng = np.random.default_rng(0)
    bmg = BatchMolGraph([make_molecule(rng, 1) for _ in range(4)])
    assert bmg.E.shape == (0, 14) and bmg.edge_index.shape == (2, 0) and bmg.batch.tolist() == [0, 1, 2, 3]


def test_integration_md_ctypes_stub_matches_the_binding():
    """INTEGRATION.md section 3 shows the ctypes stub a host language would write for the fused depth step: its argument
    list must be the one `_lib.SIGNATURES` (checked symbol for symbol against include/dmpnn.h above) binds."""
    import ctypes as C
    import os
    import re

    from chemprop_b200 import _lib

    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    m = re.search(r"lib\.dmpnn_bond_step_fused_bf16\.argtypes = (.*?)\nrc = ", doc, re.S)
    assert m, "the stub is gone from INTEGRATION.md"
    argtypes = eval(m.group(1).replace("\\\n", " "), {"C": C})       # noqa: S307 -- our own document
    res, args = _lib.SIGNATURES["dmpnn_bond_step_fused_bf16"]
    assert res is C.c_int and list(argtypes) == list(args)
    call = re.search(r"rc = lib\.dmpnn_bond_step_fused_bf16\((.*?)\)\nif rc", doc, re.S).group(1)
    n_args = len([a for a in re.sub(r"#.*", "", call).replace("\n", " ").split(",") if a.strip()])
    assert n_args == len(args), (n_args, len(args))
