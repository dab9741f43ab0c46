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

When the feature length is a multiple of 4 both stages run as the BALANCED stream kernels of three local plans
(``hg_plan_edge_reduce`` / ``hg_plan_edge_scatter``): stage A over the boundary hyperedges (side stream, then the
exchange) and over the interior hyperedges (main stream, overlapping the exchange) write one ``[interior |
boundary]`` hyperedge-feature matrix, and ONE stage B over all local hyperedges writes every row of ``Y[V_r]``
exactly once -- no zero-fill, no atomics.  The hyperedge scale ``s1*s2`` is applied to the partial rows before
they are summed (it is linear).  Other feature lengths keep the CSR kernels above.

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

    # the balanced stream stages over a local plan (feature lengths that are multiples of 4)
    def prepare_plan(self, ptr, ind, num_local, nrows):
        return self.prepare_interior(ptr, ind, num_local, nrows)

    def plan_reduce(self, plan, X, scale, a_in, out):
        from . import ops
        return ops.edge_reduce(plan, X, s1=scale, a_in=a_in, out=out)

    def plan_scatter(self, plan, Xe, a_out, out):
        from . import ops
        return ops.edge_scatter(plan, Xe, a_out=a_out, out=out)

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
    block and the hyperedge scales ``s1``, ``s2`` given as GLOBAL ``[M]`` arrays.

    The boundary hyperedges of this rank are re-ordered once as ``[owned | owned by rank 0 | rank 1 | ...]`` so
    that every block of the exchange is a CONTIGUOUS slice of the partial-feature matrix ``P``: the partial rows
    go out straight from ``P`` (no gather), the completed rows come back straight into ``P`` (no scatter), and
    ``P`` itself is what stage 2 reads.  Per call, the only feature-wide copies left beside the two kernels and
    the two ``all_to_all`` are the accumulation of the received partial rows into the owned rows and the gather of
    the completed rows each peer asked for.  All exchange buffers are allocated once per feature length."""

    def __init__(self, info: PartitionInfo, backend, group=None, split_stage_a: bool = True):
        # split_stage_a (balanced path): stage A of the boundary hyperedges runs first on a side stream so that the
        # exchange overlaps stage A of the interior ones; False = ONE stage A over every local hyperedge (each X row is
        # gathered by one kernel: better L2 re-use, nothing overlaps the exchange)
        self.info, self.backend, self.group, self.split_stage_a = info, backend, group, split_stage_a
        self.plan = backend.prepare_interior(info.int_ptr, info.int_ind, info.num_local, info.int_edges.numel())
        dev = info.bnd_ptr.device
        world, rank = info.world, info.rank
        # ---- boundary rows in exchange order
        peers = [q for q in range(world) if q != rank]
        order = torch.cat([info.own_rows.to(torch.int64)] + [info.send_rows[q].to(torch.int64) for q in peers]) \
            if info.bnd_edges.numel() else torch.empty(0, dtype=torch.int64, device=dev)
        lens = (info.bnd_ptr[1:] - info.bnd_ptr[:-1]).to(torch.int64)
        new_len = lens[order]
        self.bnd_ptr = torch.zeros(order.numel() + 1, dtype=torch.int32, device=dev)
        if order.numel():
            self.bnd_ptr[1:] = torch.cumsum(new_len, 0).to(torch.int32)
            starts = info.bnd_ptr[:-1].to(torch.int64)[order]
            rep = torch.repeat_interleave(torch.arange(order.numel(), device=dev), new_len)
            within = torch.arange(int(new_len.sum()), device=dev) - self.bnd_ptr[:-1].to(torch.int64)[rep]
            self.bnd_ind = info.bnd_ind[(starts[rep] + within)].contiguous()
        else:
            self.bnd_ind = info.bnd_ind
        self.bnd_edges = info.bnd_edges[order] if order.numel() else info.bnd_edges
        self.n_own = int(info.own_rows.numel())
        # send block of peer q = rows [send_off[q], send_off[q] + send_counts[q]) of P (rank's own block is empty)
        self.send_counts = [0 if q == rank else int(info.send_rows[q].numel()) for q in range(world)]
        self.recv_counts = [int(t.numel()) for t in info.recv_own_pos]
        nrecv = sum(self.recv_counts)
        self.recv_map = torch.cat(info.recv_own_pos).to(torch.int32) if nrecv else torch.empty(0, dtype=torch.int32, device=dev)
        self.recv_map64 = self.recv_map.to(torch.int64)
        self.recv_ptr = torch.arange(nrecv + 1, dtype=torch.int32, device=dev)
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        # local plans for the balanced stages: boundary hyperedges alone (stage A, side stream) and every local
        # hyperedge in the order [interior | boundary] (stage B); the interior plan doubles as stage A of the interior
        self.n_int, self.n_bnd = int(info.int_edges.numel()), int(self.bnd_edges.numel())
        self.plan_bnd = self.plan_all = None
        if hasattr(backend, "plan_reduce") and self.n_int + self.n_bnd > 0:
            if self.n_bnd:
                self.plan_bnd = backend.prepare_plan(self.bnd_ptr, self.bnd_ind, info.num_local, self.n_bnd)
            ptr_all = torch.cat([info.int_ptr.to(torch.int64), int(info.int_ptr[-1]) + self.bnd_ptr[1:].to(torch.int64)]).to(torch.int32)
            ind_all = torch.cat([info.int_ind, self.bnd_ind]).contiguous()
            self.plan_all = backend.prepare_plan(ptr_all, ind_all, info.num_local, self.n_int + self.n_bnd)
        self._xe = {}            # F -> [n_int + n_bnd, F] hyperedge features of the balanced path
        self.bytes_exchanged = 0
        self._bufs = {}          # F -> (recv, ret): exchange buffers, allocated once
        self._scales = {}        # cached per-hyperedge scale gathers (identity of s1 / s2 + version)

    def _buffers(self, F, like):
        b = self._bufs.get(F)
        if b is None:
            nrecv = sum(self.recv_counts)
            b = (torch.empty((nrecv, F), dtype=like.dtype, device=like.device),
                 torch.empty((nrecv, F), dtype=like.dtype, device=like.device))
            self._bufs[F] = b
        return b

    def _a2a(self, out, inp, out_counts, in_counts, F):
        import torch.distributed as dist
        if self.info.world > 1:
            dist.all_to_all_single(out, inp, output_split_sizes=out_counts, input_split_sizes=in_counts, group=self.group)
        self.bytes_exchanged += 4 * F * (sum(out_counts) + sum(in_counts))

    def _edge_scales(self, s1, s2):
        """scale of the interior and of the boundary hyperedges; gathered once per (s1, s2) contents"""
        key = tuple((t.data_ptr(), t._version) if t is not None else None for t in (s1, s2))
        hit = self._scales.get("key")
        if hit == key:
            return self._scales["int1"], self._scales["int2"], self._scales["bnd"]
        info = self.info
        pick = lambda s, e: None if s is None else s[e].contiguous()
        i1, i2 = pick(s1, info.int_edges), pick(s2, info.int_edges)
        bnd = None
        if s1 is not None or s2 is not None:
            bnd = (pick(s1, self.bnd_edges) if s1 is not None else 1.0) * (pick(s2, self.bnd_edges) if s2 is not None else 1.0)
            bnd = bnd.contiguous()
        self._scales = {"key": key, "int1": i1, "int2": i2, "bnd": bnd,
                        "int": None if (i1 is None and i2 is None) else ((i1 if i1 is not None else 1.0) * (i2 if i2 is not None else 1.0)).contiguous()}
        return i1, i2, bnd

    def forward(self, X, s1=None, s2=None, a_out=None, a_in=None):
        info, be = self.info, self.backend
        if X.shape[0] != info.num_local:
            raise ValueError(f"X has {X.shape[0]} rows, this rank owns {info.num_local} vertices")
        F = X.shape[1]
        flat = lambda t: None if t is None else t.reshape(-1).contiguous()
        s1, s2, a_out, a_in = flat(s1), flat(s2), flat(a_out), flat(a_in)
        si1, si2, sbnd = self._edge_scales(s1, s2)
        Y = torch.empty_like(X)
        use_side = self.side is not None
        if self.plan_all is not None and F % 4 == 0 and F > 0:
            return self._forward_planned(X, Y, self._scales["int"], sbnd, a_out, a_in, use_side)
        if use_side:
            self.side.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(self.side) if use_side else _NullCtx()
        with ctx:
            # boundary stage 1 + exchange (side stream), overlapping the interior kernel below
            P = be.edge_reduce(self.bnd_ptr, self.bnd_ind, X, a_in)                      # [own | to rank 0 | 1 | ...]
            if info.world > 1:
                recv, ret = self._buffers(F, P)
                n_own = self.n_own
                self._a2a(recv, P[n_own:], self.recv_counts, self.send_counts, F)         # partial rows -> owners
                own = P[:n_own]
                be.edge_scatter(self.recv_ptr, self.recv_map, recv, None, None, own)      # own[map[i]] += recv[i]
                torch.index_select(own, 0, self.recv_map64, out=ret)                      # completed rows each peer asked for
                self._a2a(P[n_own:], ret, self.send_counts, self.recv_counts, F)          # ... straight back into P
        # interior hyperedges: the single-GPU kernels on the local sub-hypergraph
        be.interior(self.plan, X, si1, si2, a_out, a_in, Y)
        if use_side:
            torch.cuda.current_stream().wait_stream(self.side)
        be.edge_scatter(self.bnd_ptr, self.bnd_ind, P, sbnd, a_out, Y)                    # boundary stage 2
        if use_side:
            P.record_stream(torch.cuda.current_stream())
        return Y


    def _forward_planned(self, X, Y, sint, sbnd, a_out, a_in, use_side):
        """Both stages as balanced stream kernels over local plans (module docstring)."""
        info, be = self.info, self.backend
        F, n_int, n_own = X.shape[1], self.n_int, self.n_own
        Xe = self._xe.get(F)
        if Xe is None:
            Xe = self._xe[F] = torch.empty((n_int + self.n_bnd, F), dtype=X.dtype, device=X.device)
        P = Xe[n_int:]                                           # [own | to rank 0 | 1 | ...]: boundary rows
        if not self.split_stage_a:
            sall = self._scales.get("all")
            if sall is None and (sint is not None or sbnd is not None):
                one = lambda n: torch.ones(n, dtype=X.dtype, device=X.device)
                sall = self._scales["all"] = torch.cat([sint if sint is not None else one(n_int),
                                                        sbnd if sbnd is not None else one(self.n_bnd)]).contiguous()
            be.plan_reduce(self.plan_all, X, sall, a_in, Xe)
            if info.world > 1 and self.n_bnd:
                recv, ret = self._buffers(F, P)
                self._a2a(recv, P[n_own:], self.recv_counts, self.send_counts, F)
                own = P[:n_own]
                be.edge_scatter(self.recv_ptr, self.recv_map, recv, None, None, own)
                torch.index_select(own, 0, self.recv_map64, out=ret)
                self._a2a(P[n_own:], ret, self.send_counts, self.recv_counts, F)
            be.plan_scatter(self.plan_all, Xe, a_out, Y)
            return Y
        if use_side:
            self.side.wait_stream(torch.cuda.current_stream())
        with (torch.cuda.stream(self.side) if use_side else _NullCtx()):
            if self.plan_bnd is not None:
                be.plan_reduce(self.plan_bnd, X, sbnd, a_in, P)                           # scaled partial rows
                if info.world > 1:
                    recv, ret = self._buffers(F, P)
                    self._a2a(recv, P[n_own:], self.recv_counts, self.send_counts, F)     # partial rows -> owners
                    own = P[:n_own]
                    be.edge_scatter(self.recv_ptr, self.recv_map, recv, None, None, own)  # own[map[i]] += recv[i]
                    torch.index_select(own, 0, self.recv_map64, out=ret)                  # completed rows each peer asked for
                    self._a2a(P[n_own:], ret, self.send_counts, self.recv_counts, F)      # ... straight back into P
        if self.plan is not None:
            be.plan_reduce(self.plan, X, sint, a_in, Xe[:n_int])                          # interior rows, overlapping the exchange
        if use_side:
            torch.cuda.current_stream().wait_stream(self.side)
        be.plan_scatter(self.plan_all, Xe, a_out, Y)                                      # every row of Y written once
        return Y


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
