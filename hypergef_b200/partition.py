"""Vertex / hyperedge PARTITIONED aggregation across GPUs (option 2).

The reference is single-GPU; this is the new capability BASELINE.json config 5 asks for.
One process per GPU (``torch.distributed``, NCCL over NVLink).  Rank ``r`` owns a contiguous
vertex block ``V_r`` and the rows ``X[V_r]``, ``Y[V_r]``.  A hyperedge whose members all lie in one
block is INTERIOR to that rank and goes through the single-GPU fused kernel untouched.  The other
(BOUNDARY) hyperedges are completed across ranks:

    stage 1   P_r[e] = sum_{u in e, u in V_r} a_in[u] X[u]              (hg_edge_reduce, local)
    exchange  partial rows -> the hyperedge's owner (rank of its lowest member), summed there
              (hg_edge_scatter with an identity CSR), completed rows -> back to every rank that
              touches the hyperedge: two all_to_all's that carry boundary rows only
    stage 2   Y[v] += a_out[v] * s[e] * Xe[e]   for v in e, v in V_r     (hg_edge_scatter, local)

so no rank ever issues a remote atomic and the payload is ``~2 * 4F * E_boundary`` bytes per rank.
Boundary stage 1 and the exchange run on a side stream while the interior kernel computes.

``build_partition`` is pure index arithmetic in torch (runs on CPU or GPU, tested with gloo at
world size 2); the compute goes through a small backend interface whose product implementation
(:class:`CudaBackend`) is the C-ABI library -- there is no CPU compute path in this package.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _native

__all__ = ["PartitionInfo", "build_partition", "PartitionedAggregator", "CudaBackend", "vertex_blocks"]


def vertex_blocks(num_nodes: int, world: int) -> List[int]:
    """Block boundaries ``b[0..world]``: rank r owns vertices ``[b[r], b[r+1])``."""
    return [(num_nodes * r) // world for r in range(world + 1)]


@dataclass
class PartitionInfo:
    rank: int
    world: int
    v_start: int
    v_end: int
    num_edges: int                      # global hyperedge count
    # interior hyperedges of this rank: CSR over LOCAL vertex ids, and their global ids
    int_ptr: torch.Tensor
    int_ind: torch.Tensor
    int_edges: torch.Tensor
    # boundary hyperedges touching this rank (ascending global id), restricted to local members
    bnd_ptr: torch.Tensor
    bnd_ind: torch.Tensor
    bnd_edges: torch.Tensor
    bnd_owner: torch.Tensor             # owning rank of each boundary hyperedge
    # exchange lists (rows are positions in bnd_edges); all ascending in global hyperedge id
    send_rows: List[torch.Tensor]       # to owner q: partial rows of hyperedges owned by q != rank
    own_rows: torch.Tensor              # positions of the hyperedges this rank owns
    recv_own_pos: List[torch.Tensor]    # from peer p: position in own_rows of each received row
    num_boundary_total: int             # global count of boundary hyperedges (for reporting)

    @property
    def num_local(self) -> int:
        return self.v_end - self.v_start


def build_partition(H_T_csrptr: torch.Tensor, H_T_colind: torch.Tensor, num_nodes: int, num_edges: int,
                    world: int, rank: int) -> PartitionInfo:
    """Partition the global ``H^T`` CSR for ``rank`` of ``world`` (every rank runs this on the same
    replicated index arrays; features are never replicated)."""
    dev = H_T_colind.device
    ptr = H_T_csrptr.to(torch.int64)
    col = H_T_colind.to(torch.int64)
    M = int(num_edges)
    bounds = torch.tensor(vertex_blocks(num_nodes, world), device=dev, dtype=torch.int64)
    deg = ptr[1:] - ptr[:-1]
    eid = torch.repeat_interleave(torch.arange(M, device=dev, dtype=torch.int64), deg)
    rk = torch.bucketize(col, bounds[1:], right=True)                 # rank of every member
    big = torch.full((M,), world, device=dev, dtype=torch.int64)
    rmin = big.scatter_reduce(0, eid, rk, reduce="amin", include_self=True)
    rmax = torch.full((M,), -1, device=dev, dtype=torch.int64).scatter_reduce(0, eid, rk, reduce="amax",
                                                                             include_self=True)
    nonempty = deg > 0
    interior = nonempty & (rmin == rmax)
    boundary = nonempty & (rmin != rmax)
    owner = rmin                                                     # colind ascending: first member's rank
    # (hyperedge, rank) incidence of the boundary hyperedges, unique and sorted by (rank, hyperedge)
    bsel = boundary[eid]
    pairs = torch.unique(rk[bsel] * M + eid[bsel])
    pair_rank, pair_edge = pairs // M, pairs % M

    v0, v1 = int(bounds[rank]), int(bounds[rank + 1])
    local = rk == rank

    def sub_csr(edge_mask):
        sel = local & edge_mask[eid]
        e_sel, v_sel = eid[sel], col[sel] - v0
        edges, inv = torch.unique(e_sel, return_inverse=True)        # ascending global ids
        cnt = torch.bincount(inv, minlength=edges.numel())
        p = torch.zeros(edges.numel() + 1, device=dev, dtype=torch.int64)
        p[1:] = torch.cumsum(cnt, 0)
        # e_sel is already grouped by hyperedge in ascending order (CSR order), members ascending
        return p.to(torch.int32), v_sel.to(torch.int32), edges

    int_ptr, int_ind, int_edges = sub_csr(interior & (rmin == rank))
    bnd_ptr, bnd_ind, bnd_edges = sub_csr(boundary)
    bnd_owner = owner[bnd_edges]
    pos_of = torch.full((M,), -1, device=dev, dtype=torch.int64)
    pos_of[bnd_edges] = torch.arange(bnd_edges.numel(), device=dev)
    send_rows = [torch.nonzero(bnd_owner == q).flatten() if q != rank else
                 torch.empty(0, device=dev, dtype=torch.int64) for q in range(world)]
    own_rows = torch.nonzero(bnd_owner == rank).flatten()
    own_pos = torch.full((M,), -1, device=dev, dtype=torch.int64)
    own_pos[bnd_edges[own_rows]] = torch.arange(own_rows.numel(), device=dev)
    recv_own_pos = []
    for p in range(world):
        if p == rank:
            recv_own_pos.append(torch.empty(0, device=dev, dtype=torch.int64))
            continue
        e_p = pair_edge[pair_rank == p]                              # boundary hyperedges touching p, ascending
        recv_own_pos.append(own_pos[e_p[owner[e_p] == rank]])
    return PartitionInfo(rank, world, v0, v1, M, int_ptr, int_ind, int_edges, bnd_ptr, bnd_ind, bnd_edges,
                         bnd_owner, send_rows, own_rows, recv_own_pos, int(boundary.sum()))


class CudaBackend:
    """The product compute backend: everything through libhgef_b200.so."""

    def __init__(self, device: torch.device, ngs: int):
        self.device, self.ngs = torch.device(device), int(ngs)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()

    def _stream(self):
        return torch.cuda.current_stream(self.index).cuda_stream

    def prepare_interior(self, ptr, ind, num_local, num_int_edges):
        from .balancer import balance_schedule
        from . import ops
        if num_int_edges == 0 or ind.numel() == 0:
            return None
        bs = balance_schedule(self.ngs, ptr)
        return ops.Plan(bs.balan_key, bs.balan_row, bs.group_st, bs.group_ed, ind.contiguous(), num_local,
                        num_int_edges)

    def interior(self, plan, X, s1, s2, a_out, a_in, out):
        from . import ops
        if plan is None:
            return out.zero_()
        return ops.aggregate(plan, X, s1=s1, s2=s2, a_out=a_out, a_in=a_in, out=out)

    def edge_reduce(self, ptr, ind, X, a_in):
        nrow, F = ptr.numel() - 1, X.shape[1]
        P = torch.empty((nrow, F), dtype=torch.float32, device=X.device)
        if nrow:
            _native.call("hg_edge_reduce", nrow, ptr.data_ptr(), ind.data_ptr(), X.data_ptr(),
                         None if a_in is None else a_in.data_ptr(), P.data_ptr(), F, self.index, self._stream())
        return P

    def edge_scatter(self, ptr, ind, Q, scale, a_out, Y):
        nrow, F = ptr.numel() - 1, Q.shape[1]
        if nrow:
            _native.call("hg_edge_scatter", nrow, ptr.data_ptr(), ind.data_ptr(), Q.data_ptr(),
                         None if scale is None else scale.data_ptr(),
                         None if a_out is None else a_out.data_ptr(), Y.data_ptr(), F, self.index,
                         self._stream())
        return Y


class PartitionedAggregator:
    """``Y[V_r] = (diag(a_out) H diag(s1*s2) H^T diag(a_in) X)[V_r]`` with X, Y, a_* sharded by vertex
    block and the hyperedge scales ``s1``, ``s2`` given as GLOBAL ``[M]`` arrays."""

    def __init__(self, info: PartitionInfo, backend, group=None):
        self.info, self.backend, self.group = info, backend, group
        self.plan = backend.prepare_interior(info.int_ptr, info.int_ind, info.num_local, info.int_edges.numel())
        dev = info.bnd_ptr.device
        nrecv = sum(int(t.numel()) for t in info.recv_own_pos)
        self.recv_map = torch.cat(info.recv_own_pos).to(torch.int32) if nrecv else torch.empty(0, dtype=torch.int32, device=dev)
        self.recv_ptr = torch.arange(nrecv + 1, dtype=torch.int32, device=dev)
        self.send_cat = torch.cat(info.send_rows) if info.world > 1 else torch.empty(0, dtype=torch.int64, device=dev)
        self.send_counts = [int(t.numel()) for t in info.send_rows]
        self.recv_counts = [int(t.numel()) for t in info.recv_own_pos]
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.bytes_exchanged = 0

    def _all_to_all(self, send, send_counts, recv_counts, F):
        import torch.distributed as dist
        recv = torch.empty((sum(recv_counts), F), dtype=send.dtype, device=send.device)
        if self.info.world > 1:
            dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=recv_counts,
                                   input_split_sizes=send_counts, group=self.group)
        self.bytes_exchanged += 4 * F * (sum(send_counts) + sum(recv_counts))
        return recv

    def forward(self, X, s1=None, s2=None, a_out=None, a_in=None):
        info, be = self.info, self.backend
        if X.shape[0] != info.num_local:
            raise ValueError(f"X has {X.shape[0]} rows, this rank owns {info.num_local} vertices")
        F = X.shape[1]
        flat = lambda t: None if t is None else t.reshape(-1).contiguous()
        s1, s2, a_out, a_in = flat(s1), flat(s2), flat(a_out), flat(a_in)
        pick = lambda s, e: None if s is None else s[e].contiguous()
        Y = torch.empty_like(X)
        use_side = self.side is not None
        if use_side:
            self.side.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(self.side) if use_side else _NullCtx()
        with ctx:
            # boundary stage 1 + exchange (side stream), overlapping the interior kernel below
            P = be.edge_reduce(info.bnd_ptr, info.bnd_ind, X, a_in)                    # [B_r, F]
            Q = P                                                                         # completed in place
            if info.world > 1:
                recv = self._all_to_all(P[self.send_cat], self.send_counts, self.recv_counts, F)
                own = P[info.own_rows].contiguous()                                      # [O_r, F]
                be.edge_scatter(self.recv_ptr, self.recv_map, recv, None, None, own)     # own[map[i]] += recv[i]
                back = self._all_to_all(own[self.recv_map.long()], self.recv_counts, self.send_counts, F)
                Q = P.clone()
                Q[info.own_rows] = own
                Q[self.send_cat] = back
        # interior hyperedges: the single-GPU fused kernel on the local sub-hypergraph
        be.interior(self.plan, X, pick(s1, info.int_edges), pick(s2, info.int_edges), a_out, a_in, Y)
        if use_side:
            torch.cuda.current_stream().wait_stream(self.side)
        scale = None
        if s1 is not None or s2 is not None:
            scale = (pick(s1, info.bnd_edges) if s1 is not None else 1.0) * (pick(s2, info.bnd_edges) if s2 is not None else 1.0)
            scale = scale.contiguous()
        be.edge_scatter(info.bnd_ptr, info.bnd_ind, Q, scale, a_out, Y)                  # boundary stage 2
        if use_side:
            for t in (P, Q):
                t.record_stream(torch.cuda.current_stream())
        return Y


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
