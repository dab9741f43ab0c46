"""On-disk incidence formats.

``read_mtx`` replaces the reference's C++ MatrixMarket loader (``include/dataloader/dataloader.hpp:22-104``,
used by its standalone binaries; ``HyperGraph.store_mtx`` is the writer, ``hypergraph.py:79-85``): coordinate
format, values dropped, 0-based, ``symmetric`` mirrored and de-duplicated, coordinates sorted row-major.
Rows are vertices, columns are hyperedges.  ``hypergraph_from_mtx`` builds the ``edge_index`` layout the
reference's ``HyperGraph`` consumes (``[[V ; E+N], [E+N ; V]]``, ``hypergraph.py:15``) and hands it over.

``load_citation_dataset`` / ``load_cornell_dataset`` / ``load_LE_dataset`` read the three raw on-disk layouts of the
AllSet collection the reference trains on (``data/load_dataset.py:33-119,122-196,294-384``; selected by name in
``prepare_data.py``) into the same ``data`` object, so the real datasets can be dropped in where the synthetic
shapes are used.  They need only numpy / scipy / torch (the reference's versions need torch_geometric's ``Data``
and torch_sparse's ``coalesce``; the latter is restated here as "sort the (row, col) pairs, drop duplicates").
The files themselves are not in the reference checkout and cannot be fetched here: the readers are tested on
small files written in the same layouts (``tests/test_io.py``), not against the published datasets.
"""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace

import torch

from . import _native

__all__ = ["read_mtx", "data_from_mtx", "hypergraph_from_mtx", "load_citation_dataset", "load_cornell_dataset",
           "load_LE_dataset", "data_from_members"]


def read_mtx(path):
    """``(V, E, num_nodes, num_edges)``: int64 CPU tensors of the sorted incidence pairs of ``path``."""
    handle, nrow, ncol, nnz = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
    _native.call("hg_mtx_open", os.fsencode(path), C.byref(handle), C.byref(nrow), C.byref(ncol), C.byref(nnz))
    try:
        V = torch.empty(nnz.value, dtype=torch.int64)
        E = torch.empty(nnz.value, dtype=torch.int64)
        _native.call("hg_mtx_fill", handle, V.data_ptr(), E.data_ptr())
    finally:
        _native.call("hg_mtx_close", handle)
    return V, E, nrow.value, ncol.value


def data_from_mtx(path, num_feat: int = 0, seed: int = 0):
    """A ``data`` object (``x``, ``edge_index``) as ``HyperGraph.__init__`` expects it.  Hyperedges with no
    member are dropped and the rest renumbered consecutively, which is what the reference's constructor
    requires (``num_edges = len(unique(E))``, ``hypergraph.py:19``)."""
    V, E, N, M = read_mtx(path)
    used, E = torch.unique(E, return_inverse=True)
    M = int(used.numel())
    order = torch.argsort(E * N + V)          # second half sorted by hyperedge: its first row id must be N (:15)
    ei = torch.stack([torch.cat([V, E[order] + N]), torch.cat([E + N, V[order]])])
    x = torch.randn(N, num_feat, generator=torch.Generator().manual_seed(seed)) if num_feat else torch.zeros(N, 1)
    return SimpleNamespace(x=x, edge_index=ei, num_nodes=N, num_hyperedges=M, nnz=int(V.numel()))


def hypergraph_from_mtx(path, device, ngs, data_name="mtx"):
    from .hypergraph import HyperGraph
    return HyperGraph(data_from_mtx(path), device, data_name, ngs=ngs)


# ---------------------------------------------------------------------------------------------------
# raw AllSet layouts -> data (x, y, edge_index, n_x, num_hyperedges)
# ---------------------------------------------------------------------------------------------------
def data_from_members(node_ids, he_ids, num_nodes, x, y, train_percent=0.025):
    """``node_ids[i]`` is a member of hyperedge ``he_ids[i]`` (hyperedges numbered from 0).  Builds the bipartite
    ``edge_index = [[V ; E+N], [E+N ; V]]`` with the (row, col) pairs sorted and duplicates dropped -- what
    ``torch_sparse.coalesce`` does in ``data/load_dataset.py:168-171`` -- so the first ``nnz`` columns are the
    vertex -> hyperedge half that ``HyperGraph.__init__`` cuts at the first row id ``>= N`` (``hypergraph.py:15``)."""
    V = torch.as_tensor(node_ids, dtype=torch.int64).reshape(-1)
    E = torch.as_tensor(he_ids, dtype=torch.int64).reshape(-1)
    if V.numel() != E.numel():
        raise ValueError("node and hyperedge lists differ in length")
    N = int(num_nodes)
    if V.numel() and (int(V.min()) < 0 or int(V.max()) >= N or int(E.min()) < 0):
        raise ValueError("member ids out of range")
    M = int(E.max()) + 1 if E.numel() else 0
    total = N + M
    rows = torch.cat([V, E + N])
    cols = torch.cat([E + N, V])
    key = torch.unique(rows * total + cols)                 # sorted, duplicates removed
    ei = torch.stack([key // total, key % total])
    return SimpleNamespace(x=x, y=y, edge_index=ei, n_x=N, num_nodes=N, num_hyperedges=M,
                           nnz=int(key.numel()) // 2, train_percent=train_percent)


def load_citation_dataset(path, dataset="cora", train_percent=0.025):
    """HyperGCN's pickles (``data/load_dataset.py:122-196``): ``features.pickle`` (a scipy sparse matrix),
    ``labels.pickle``, ``hypergraph.pickle`` (``{hyperedge: [member vertices]}``; hyperedges are numbered in the
    dictionary's order)."""
    import pickle
    import numpy as np
    root = os.path.join(path, dataset)
    with open(os.path.join(root, "features.pickle"), "rb") as f:
        feats = pickle.load(f)
    feats = np.asarray(feats.todense() if hasattr(feats, "todense") else feats, dtype=np.float32)
    with open(os.path.join(root, "labels.pickle"), "rb") as f:
        labels = np.asarray(pickle.load(f))
    if feats.shape[0] != labels.shape[0]:
        raise ValueError(f"{feats.shape[0]} feature rows for {labels.shape[0]} labels")
    with open(os.path.join(root, "hypergraph.pickle"), "rb") as f:
        hyper = pickle.load(f)
    nodes, hes = [], []
    for k, members in enumerate(hyper.values()):
        members = list(members)
        nodes += members
        hes += [k] * len(members)
    data = data_from_members(nodes, hes, feats.shape[0], torch.from_numpy(feats), torch.from_numpy(labels).long(),
                             train_percent)
    data.num_hyperedges = len(hyper)
    return data


def load_cornell_dataset(path, dataset="walmart-trips", feature_noise=0.1, feature_dim=None, train_percent=0.025,
                         seed=None):
    """Cornell text files (``data/load_dataset.py:294-384``): ``node-labels-<name>.txt`` (one label per line, from 1),
    ``hyperedges-<name>.txt`` (one hyperedge per line, comma-separated vertex ids, shifted so that the smallest id
    is 0).  Features are the one-hot label (zero-padded to ``feature_dim``) plus Gaussian noise of width
    ``feature_noise``, as the reference draws them (``seed`` makes the draw repeatable; the reference does not seed)."""
    import numpy as np
    root = os.path.join(path, dataset)
    labels = np.loadtxt(os.path.join(root, f"node-labels-{dataset}.txt"), dtype=np.int64).reshape(-1)
    N, ncls = labels.shape[0], int(labels.max())
    width = max(ncls, feature_dim or 0)
    onehot = np.zeros((N, width), dtype=np.float64)
    onehot[np.arange(N), labels - 1] = 1.0
    feats = np.random.default_rng(seed).normal(onehot, feature_noise)
    nodes, hes = [], []
    with open(os.path.join(root, f"hyperedges-{dataset}.txt")) as f:
        k = 0
        for line in f:
            line = line.strip()
            if not line:
                continue
            members = [int(t) for t in line.split(",")]
            nodes += members
            hes += [k] * len(members)
            k += 1
    lo = min(nodes) if nodes else 0
    data = data_from_members([v - lo for v in nodes], hes, N, torch.from_numpy(feats).float(),
                             torch.from_numpy(labels).long(), train_percent)
    data.num_hyperedges = k
    return data


def load_LE_dataset(path, dataset="ModelNet40", train_percent=0.025):
    """The ``.content`` / ``.edges`` pair (``data/load_dataset.py:33-119``): ``<name>.content`` rows are
    ``id  features...  label``; ``<name>.edges`` rows are ``(vertex id, hyperedge id)`` in ONE id space in which
    every hyperedge id is above every vertex id; ids are renumbered in the order of the ``.content`` rows."""
    import numpy as np
    root = os.path.join(path, dataset)
    content = np.genfromtxt(os.path.join(root, f"{dataset}.content"), dtype=str)
    if content.ndim == 1:
        content = content[None, :]
    ids = content[:, 0].astype(np.int64)
    feats = content[:, 1:-1].astype(np.float32)
    labels = content[:, -1].astype(np.float64).astype(np.int64)
    remap = {int(j): i for i, j in enumerate(ids)}
    edges = np.genfromtxt(os.path.join(root, f"{dataset}.edges"), dtype=np.int64).reshape(-1, 2)
    pairs = np.array([[remap[int(a)], remap[int(b)]] for a, b in edges], dtype=np.int64)
    N = int(pairs[:, 0].max()) + 1
    if int(pairs[:, 1].min()) != N:
        raise ValueError("hyperedge ids must start right after the vertex ids (data/load_dataset.py:71)")
    if np.unique(pairs).size != int(pairs.max()) + 1:
        raise ValueError("vertex / hyperedge ids are not consecutive (data/load_dataset.py:74)")
    # (the .content file also lists the hyperedge ids; only the vertex rows are kept, data/load_dataset.py:84-86)
    return data_from_members(pairs[:, 0], pairs[:, 1] - N, N, torch.from_numpy(feats[:N].copy()),
                             torch.from_numpy(labels[:N].copy()), train_percent)
