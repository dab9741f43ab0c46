"""MatrixMarket incidence reader (host) against scipy.io and, where oracle/_ref was built, the
reference's own read_mtx_file compiled in place.  Golden files live in tests/golden/ (hand-written, tiny)."""
import os

import numpy as np
import pytest
import torch

import hypergef_b200 as hgef
from hypergef_b200 import _native, io as hio
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_golden_pattern_file():
    V, E, N, M = hio.read_mtx(os.path.join(GOLD, "incidence_toy.mtx"))
    assert (N, M) == (6, 4)
    # rows = vertices, columns = hyperedges, sorted row-major, 0-based
    assert V.tolist() == [0, 0, 1, 2, 2, 3, 4, 4, 5]
    assert E.tolist() == [0, 2, 0, 1, 2, 1, 2, 3, 3]
    data = hio.data_from_mtx(os.path.join(GOLD, "incidence_toy.mtx"))
    assert data.edge_index.shape == (2, 18) and int(data.edge_index[0, 9]) == 6   # E + N half starts at column nnz


@pytest.mark.parametrize("field", ["pattern", "real", "integer"])
def test_against_scipy_and_reference(tmp_path, field):
    import scipy.sparse as sp
    from scipy.io import mmread, mmwrite
    rng = np.random.default_rng(7)
    N, M, Z = 300, 120, 900
    r, c = rng.integers(0, N, Z), rng.integers(0, M, Z)
    keep = np.unique(np.stack([r, c]), axis=1)
    vals = np.ones(keep.shape[1]) if field != "real" else rng.random(keep.shape[1]) + 0.5
    A = sp.coo_matrix((vals if field != "integer" else vals.astype(np.int64), (keep[0], keep[1])), shape=(N, M))
    path = str(tmp_path / f"rand_{field}.mtx")
    mmwrite(path, A, field=field, comment="a comment line\nanother")
    V, E, n, m = hio.read_mtx(path)
    assert (n, m) == (N, M)
    B = mmread(path).tocoo()
    order = np.lexsort((B.col, B.row))
    assert np.array_equal(V.numpy(), B.row[order]) and np.array_equal(E.numpy(), B.col[order])
    if orc.ref_available():
        rn, rm, indptr, indices, rowind = orc.ref_read_mtx(path)
        assert (rn, rm) == (N, M)
        assert np.array_equal(rowind, V.numpy()) and np.array_equal(indices, E.numpy())
        assert np.array_equal(indptr, np.concatenate([[0], np.cumsum(np.bincount(V.numpy(), minlength=N))]))


def test_symmetric_is_mirrored_and_deduplicated(tmp_path):
    path = _write(tmp_path, "sym.mtx", "%%MatrixMarket matrix coordinate pattern symmetric\n% c\n4 4 4\n2 1\n3 3\n4 2\n4 1\n")
    V, E, n, m = hio.read_mtx(path)
    pairs = sorted(zip(V.tolist(), E.tolist()))
    assert pairs == [(0, 1), (0, 3), (1, 0), (1, 3), (2, 2), (3, 0), (3, 1)]
    if orc.ref_available():
        # The reference mirrors and de-duplicates too, but then walks the coordinates with the FILE's entry
        # count as the bound (dataloader.hpp:92 `curr_pos < nnz`), so a symmetric file comes back truncated to
        # its first `entries` sorted pairs.  Pinned here as observed; the product returns all of them.
        _, _, _, indices, rowind = orc.ref_read_mtx(path)
        assert list(zip(rowind.tolist(), indices.tolist())) == pairs[:4]


def test_errors_are_return_codes(tmp_path):
    with pytest.raises(ValueError):
        hio.read_mtx(str(tmp_path / "missing.mtx"))
    with pytest.raises(ValueError):
        hio.read_mtx(_write(tmp_path, "nobanner.mtx", "3 3 1\n1 1\n"))
    with pytest.raises(ValueError):
        hio.read_mtx(_write(tmp_path, "array.mtx", "%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n"))
    with pytest.raises(ValueError):
        hio.read_mtx(_write(tmp_path, "short.mtx", "%%MatrixMarket matrix coordinate pattern general\n3 3 2\n1 1\n"))
    with pytest.raises(_native.HgefGraphError):
        hio.read_mtx(_write(tmp_path, "range.mtx", "%%MatrixMarket matrix coordinate pattern general\n3 3 1\n4 1\n"))


def test_host_graph_from_mtx_matches_direct_construction(tmp_path):
    """mtx -> HyperGraph (host CSR path) equals the graph built from the same pairs directly."""
    d = np.load(os.path.join(GOLD, "graph_mini.npz"))
    from scipy.io import mmwrite
    import scipy.sparse as sp
    N = int(d["num_nodes"])
    ei = d["edge_index"]
    half = ei.shape[1] // 2
    V, E = ei[0, :half], ei[1, :half] - N
    path = str(tmp_path / "mini.mtx")
    mmwrite(path, sp.coo_matrix((np.ones(V.size), (V, E)), shape=(N, int(E.max()) + 1)), field="pattern")
    hg = hio.hypergraph_from_mtx(path, torch.device("cpu"), int(d["ngs"]))
    assert np.array_equal(hg.H_T_csrptr.numpy(), d["H_T_csrptr"]) and np.array_equal(hg.H_T_colind.numpy(), d["H_T_colind"])
    assert np.array_equal(hg.group_key.numpy(), d["group_key"]) and np.array_equal(hg.group_row.numpy(), d["group_row"])
