"""Locality reordering: an optional, once-per-graph permutation of vertex and hyperedge ids.

The reference vendors Rabbit Order under ``include/reorder/`` but never calls it.  What a permutation can buy on
this path is L2 re-use: stage A gathers every X row once per hyperedge it belongs to and stage B every Xe row once
per member, and the repeats are served by the L2 only if they fall within its reach.
Graphs whose ids carry no locality (shuffled exports, hashed ids) lose those hits; a bandwidth-reducing order
brings them back.  The order used here is reverse Cuthill-McKee on the bipartite vertex/hyperedge graph
(``scipy.sparse.csgraph``), which numbers the members of a hyperedge -- and the hyperedges of a vertex -- close
together; it runs on the host, once, and the result is an ordinary ``data`` object for ``HyperGraph``.

    data2, vperm, eperm = reorder.reorder(data)          # new id = perm[old id]
    Y_old_order = reorder.restore_rows(Y2, vperm)        # rows back in the caller's numbering
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

__all__ = ["rcm_order", "permute_data", "reorder", "restore_rows", "mean_span"]


def _split(edge_index, num_nodes):
    """the vertex -> hyperedge half of ``[[V ; E+N], [E+N ; V]]`` (hypergraph.py:15-18)"""
    ei = edge_index.cpu()
    half = int((ei[0] < num_nodes).sum())
    return ei[0, :half].numpy(), (ei[1, :half] - num_nodes).numpy()


def rcm_order(V, E, num_nodes, num_edges):
    """``(vperm, eperm)`` with ``new id = perm[old id]``: reverse Cuthill-McKee over the bipartite graph, vertices
    and hyperedges each numbered in the order RCM visits them."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    V, E = np.asarray(V, dtype=np.int64), np.asarray(E, dtype=np.int64)
    N, M = int(num_nodes), int(num_edges)
    rows, cols = np.concatenate([V, E + N]), np.concatenate([E + N, V])
    A = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(N + M, N + M))
    visit = np.asarray(reverse_cuthill_mckee(A, symmetric_mode=True), dtype=np.int64)   # visit[k] = old id at position k
    is_v = visit < N
    vperm, eperm = np.empty(N, dtype=np.int64), np.empty(M, dtype=np.int64)
    vperm[visit[is_v]] = np.arange(N)
    eperm[visit[~is_v] - N] = np.arange(M)
    return torch.from_numpy(vperm), torch.from_numpy(eperm)


def permute_data(data, vperm, eperm):
    """The same hypergraph with vertex ``v`` renamed ``vperm[v]`` and hyperedge ``e`` renamed ``eperm[e]``;
    per-vertex tensors (``x``, ``y``) move with their vertices."""
    from .io import data_from_members
    N = int(getattr(data, "n_x", None) or getattr(data, "num_nodes"))
    V, E = _split(data.edge_index, N)
    x, y = getattr(data, "x", None), getattr(data, "y", None)

    def move(t):
        if t is None or t.shape[0] != N:
            return t
        out = torch.empty_like(t)
        out[vperm.to(t.device)] = t
        return out
    new = data_from_members(vperm[torch.from_numpy(V)], eperm[torch.from_numpy(E)], N, move(x), move(y))
    for k, v in vars(data).items():           # dataset name, ngs, shape ... travel unchanged
        if not hasattr(new, k):
            setattr(new, k, v)
    new.vertex_perm, new.edge_perm = vperm, eperm
    return new


def reorder(data, method="rcm"):
    """``(data', vperm, eperm)``: ``data`` renumbered for locality.  ``method``: ``"rcm"``."""
    if method != "rcm":
        raise ValueError(f"unknown reordering method {method!r}")
    N = int(getattr(data, "n_x", None) or getattr(data, "num_nodes"))
    V, E = _split(data.edge_index, N)
    M = int(E.max()) + 1 if E.size else 0
    vperm, eperm = rcm_order(V, E, N, M)
    return permute_data(data, vperm, eperm), vperm, eperm


def restore_rows(Y, vperm):
    """rows of a per-vertex result of the reordered graph, back in the original numbering"""
    return Y[vperm.to(Y.device)]


def mean_span(edge_index, num_nodes):
    """mean over hyperedges of (largest member id - smallest member id): a cheap locality figure"""
    V, E = _split(edge_index, num_nodes)
    M = int(E.max()) + 1 if E.size else 0
    lo = np.full(M, np.iinfo(np.int64).max)
    hi = np.full(M, -1)
    np.minimum.at(lo, E, V)
    np.maximum.at(hi, E, V)
    return float((hi - lo).mean()) if M else 0.0
