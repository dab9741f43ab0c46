"""On-disk incidence formats.

``read_mtx`` replaces the reference's C++ MatrixMarket loader (``include/dataloader/dataloader.hpp:22-104``,
used by its standalone binaries; ``HyperGraph.store_mtx`` is the writer, ``hypergraph.py:79-85``): coordinate
format, values dropped, 0-based, ``symmetric`` mirrored and de-duplicated, coordinates sorted row-major.
Rows are vertices, columns are hyperedges.  ``hypergraph_from_mtx`` builds the ``edge_index`` layout the
reference's ``HyperGraph`` consumes (``[[V ; E+N], [E+N ; V]]``, ``hypergraph.py:15``) and hands it over.
"""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace

import torch

from . import _native

__all__ = ["read_mtx", "data_from_mtx", "hypergraph_from_mtx"]


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
