"""Seeded synthetic hypergraphs in the reference's ``edge_index`` contract.

The reference never ships data: its loaders download AllSet pickles
(``HyperGsys/data/prepare.sh:1``) and turn them into a PyG ``Data`` whose
``edge_index`` is the coalesced (row-sorted, duplicate-free) ``[2, 2Z]`` int64
array ``[[V ; E+N], [E+N ; V]]`` (``HyperGsys/data/load_dataset.py:166-182``).
``HyperGraph.__init__`` only ever looks at that array and at ``data.x.shape[0]``
(``HyperGsys/hypergraph.py:14-20``).  These generators produce exactly that
contract for the five BASELINE.json shapes, with documented id-locality:

* ``global``   members uniform over all vertices (no locality at all);
* ``window``   hyperedge ``e`` draws each member from a window of ``window``
               consecutive vertex ids centred at ``e*N/E`` with probability
               ``p_local`` and uniformly otherwise;
* ``replicas`` a block-diagonal batch of ``k`` independently drawn copies of a
               base shape (the way GNN mini-batches stack graphs): replica ``r``
               owns vertices ``[r*N, (r+1)*N)`` and hyperedges ``[r*E, (r+1)*E)``.

Everything is torch so the large shapes can be drawn directly on the GPU; the
CPU generator (``device='cpu'``) is bit-reproducible for a given torch build
and is what the committed golden fixtures were drawn with.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Optional

import torch

__all__ = [
    "SHAPES", "HyperShape", "draw_sizes", "draw_incidence", "incidence_to_edge_index",
    "make_data", "make_shape",
]


@dataclass(frozen=True)
class HyperShape:
    """One BASELINE.json configuration made concrete."""
    name: str
    num_nodes: int
    num_edges: int
    ngs: int                       # hypergraph.py:74-75 partition_dict entry of the dataset it imitates
    size_dist: str                 # 'geom' | 'zipf'
    mean_size: float = 4.0
    min_size: int = 2
    max_size: int = 64
    zipf_alpha: float = 1.8
    force_max: int = 0             # plant one hyperedge of exactly this size (balancer stress)
    locality: str = "global"       # 'global' | 'window'
    window: int = 1 << 16
    p_local: float = 0.9
    dataset: str = ""              # name used for the ngs dictionary


SHAPES = {
    # C1: Cora-shaped, BASELINE.json configs[0]
    "cora": HyperShape("cora", 2708, 1579, 210, "geom", mean_size=3.0, min_size=2, max_size=5, dataset="cora"),
    # C2: Pubmed-shaped (AllSet co-citation pubmed: 19717 vertices, 7963 hyperedges, nnz 34795, max size 171)
    "pubmed": HyperShape("pubmed", 19717, 7963, 40, "geom", mean_size=4.37, min_size=2, max_size=171, dataset="pubmed"),
    # C3: DBLP-co-authorship-shaped
    "dblp": HyperShape("dblp", 41302, 22363, 80, "geom", mean_size=4.5, min_size=2, max_size=202, dataset="coauthor_dblp"),
    # C4: Walmart-trips-shaped power law, one planted giant hyperedge > 10^4
    "walmart": HyperShape("walmart", 88860, 69906, 210, "zipf", min_size=2, max_size=20000, zipf_alpha=1.8,
                          force_max=12345, dataset="walmart-trips"),
    # C5: 50M x 10M, mean size 10, window locality
    "c5": HyperShape("c5", 50_000_000, 10_000_000, 210, "geom", mean_size=10.0, min_size=2, max_size=512,
                     locality="window", window=1 << 16, p_local=0.9, dataset="walmart-trips"),
}


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def draw_sizes(shape: HyperShape, num_edges: int, gen: torch.Generator, device) -> torch.Tensor:
    """Hyperedge sizes (int64 ``[num_edges]``) of the shape's distribution."""
    u = torch.rand(num_edges, generator=gen, device=device, dtype=torch.float64)
    if shape.size_dist == "geom":
        # shifted geometric with the requested mean, clipped to [min, max]
        extra = max(shape.mean_size - shape.min_size, 1e-9)
        p = 1.0 / (1.0 + extra)
        k = torch.floor(torch.log1p(-u) / math.log1p(-p))
        sizes = shape.min_size + k.to(torch.int64)
    elif shape.size_dist == "zipf":
        # continuous inverse-CDF of a Pareto tail with exponent alpha, floored
        a = shape.zipf_alpha - 1.0
        sizes = torch.floor(shape.min_size * torch.pow(1.0 - u, -1.0 / a)).to(torch.int64)
    else:
        raise ValueError(f"unknown size_dist {shape.size_dist!r}")
    sizes = sizes.clamp_(shape.min_size, min(shape.max_size, shape.num_nodes))
    if shape.force_max:
        sizes[num_edges // 3] = min(shape.force_max, shape.num_nodes)
    return sizes


def draw_incidence(shape: HyperShape, replicas: int = 1, seed: int = 0, device="cpu"):
    """Draw the coalesced incidence pairs of ``replicas`` block-diagonal copies.

    Returns ``(V, E, N, M)``: int64 vertex and hyperedge id per non-zero, sorted
    by ``(V, E)`` and duplicate-free, with ``N`` vertices and ``M`` hyperedges.
    Members are drawn with replacement and then coalesced, so a hyperedge can
    end up slightly smaller than its drawn size; every hyperedge keeps at least
    one member, so hyperedge ids stay consecutive (hypergraph.py:19 relies on it).
    """
    device = torch.device(device)
    gen = _gen(seed, device)
    n, m = shape.num_nodes, shape.num_edges
    N, M = n * replicas, m * replicas
    sizes = draw_sizes(shape, M, gen, device)
    if replicas > 1 and shape.force_max:
        pass  # one planted giant per draw is enough; sizes already hold it once
    eid = torch.repeat_interleave(torch.arange(M, device=device, dtype=torch.int64), sizes)
    Z0 = eid.numel()
    rep = eid // m                                      # replica of each slot
    u = torch.rand(Z0, generator=gen, device=device, dtype=torch.float64)
    if shape.locality == "window":
        centre = ((eid % m).to(torch.float64) + 0.5) * (n / m)
        w = min(shape.window, n)
        lo = (centre - w / 2).clamp_(0, n - w).to(torch.int64)
        local = lo + torch.floor(u * w).to(torch.int64)
        u2 = torch.rand(Z0, generator=gen, device=device, dtype=torch.float64)
        pick = torch.rand(Z0, generator=gen, device=device, dtype=torch.float64) < shape.p_local
        vid = torch.where(pick, local, torch.floor(u2 * n).to(torch.int64))
    elif shape.locality == "global":
        vid = torch.floor(u * n).to(torch.int64)
    else:
        raise ValueError(f"unknown locality {shape.locality!r}")
    vid = vid.clamp_(0, n - 1) + rep * n
    # coalesce: sort by (V, E) and drop duplicates -- this is what
    # torch_sparse.coalesce does to edge_index in load_dataset.py:176-179
    key = torch.unique(vid * M + eid, sorted=True)
    V = key // M
    E = key - V * M
    return V, E, N, M


def incidence_to_edge_index(V: torch.Tensor, E: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """``[[V ; E+N],[E+N ; V]]`` row-sorted, as load_dataset.py:166-179 leaves it."""
    En = E + num_nodes
    first = torch.stack((V, En))                       # already sorted by (V, E)
    order = torch.argsort(En * (int(V.max()) + 1 if V.numel() else 1) + V)
    second = torch.stack((En[order], V[order]))
    return torch.cat((first, second), dim=1)


def make_data(V, E, num_nodes: int, num_feat: int = 0, num_class: int = 7, seed: int = 0,
              feat_device="cpu") -> SimpleNamespace:
    """A duck-typed stand-in for the PyG ``Data`` object HyperGraph consumes."""
    g = _gen(seed + 1, torch.device(feat_device))
    x = torch.randn(num_nodes, max(num_feat, 0), generator=g, device=feat_device, dtype=torch.float32)
    y = torch.randint(0, num_class, (num_nodes,), generator=g, device=feat_device)
    return SimpleNamespace(x=x, y=y, edge_index=incidence_to_edge_index(V, E, num_nodes),
                           num_nodes=num_nodes)


def make_shape(name: str, replicas: int = 1, seed: int = 0, num_feat: int = 0, device="cpu",
               shape: Optional[HyperShape] = None) -> SimpleNamespace:
    """Draw shape ``name`` and wrap it as ``data`` (+ ``.shape``, ``.ngs``, ``.nnz``)."""
    shp = shape or SHAPES[name]
    V, E, N, M = draw_incidence(shp, replicas=replicas, seed=seed, device=device)
    data = make_data(V, E, N, num_feat=num_feat, seed=seed, feat_device=device)
    data.shape_name, data.ngs, data.num_hyperedges, data.nnz = shp.name, shp.ngs, M, int(V.numel())
    data.dataset = shp.dataset
    return data
