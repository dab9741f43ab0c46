"""Host-side logic of the vertex/hyperedge-partitioned multi-GPU path, on CPU with gloo, world size 2 and 3.
The index arithmetic (build_partition) and the exchange (two all_to_all's) are the product code; the
local compute goes through a torch stand-in backend defined HERE (the product backend is CUDA-only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from hypergef_b200.partition import PartitionedAggregator, build_partition, vertex_blocks
from oracle import oracle as orc


class TorchStandInBackend:
    """Test-only: the three backend calls restated with index_add_ (fp64 for a tight check)."""

    @staticmethod
    def _rows(ptr):
        return torch.repeat_interleave(torch.arange(ptr.numel() - 1), (ptr[1:] - ptr[:-1]).long())

    def prepare_interior(self, ptr, ind, num_local, num_int_edges):
        return (ptr.long(), ind.long(), num_local)

    def interior(self, plan, X, s1, s2, a_out, a_in, out):
        ptr, ind, n = plan
        Xs = X if a_in is None else X * a_in[:, None]
        rows = self._rows(ptr)
        xe = torch.zeros(ptr.numel() - 1, X.shape[1], dtype=X.dtype).index_add_(0, rows, Xs[ind])
        for s in (s1, s2):
            if s is not None:
                xe = xe * s[:, None]
        y = torch.zeros_like(X).index_add_(0, ind, xe[rows])
        out.copy_(y if a_out is None else y * a_out[:, None])
        return out

    # the balanced-stage interface (hg_plan_edge_reduce / hg_plan_edge_scatter in the product backend)
    def prepare_plan(self, ptr, ind, num_local, nrows):
        return (ptr.long(), ind.long(), num_local)

    def plan_reduce(self, plan, X, scale, a_in, out):
        ptr, ind, n = plan
        Xs = X if a_in is None else X * a_in[:, None]
        xe = torch.zeros(ptr.numel() - 1, X.shape[1], dtype=X.dtype).index_add_(0, self._rows(ptr), Xs[ind])
        out.copy_(xe if scale is None else xe * scale[:, None])
        return out

    def plan_scatter(self, plan, Xe, a_out, out):
        ptr, ind, n = plan
        y = torch.zeros(n, Xe.shape[1], dtype=Xe.dtype).index_add_(0, ind, Xe[self._rows(ptr)])
        out.copy_(y if a_out is None else y * a_out[:, None])
        return out

    def edge_reduce(self, ptr, ind, X, a_in):
        Xs = X if a_in is None else X * a_in[:, None]
        return torch.zeros(ptr.numel() - 1, X.shape[1], dtype=X.dtype).index_add_(0, self._rows(ptr.long()), Xs[ind.long()])

    def edge_scatter(self, ptr, ind, Q, scale, a_out, Y):
        rows = self._rows(ptr.long())
        q = Q if scale is None else Q * scale[:, None]
        add = q[rows]
        if a_out is not None:
            add = add * a_out[ind.long()][:, None]
        Y.index_add_(0, ind.long(), add)
        return Y


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, graph, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = np.load(graph)
        N, M = int(d["num_nodes"]), int(d["num_edges"])
        ptr, col = torch.from_numpy(d["H_T_csrptr"]), torch.from_numpy(d["H_T_colind"])
        info = build_partition(ptr, col, N, M, world, rank)
        agg = PartitionedAggregator(info, TorchStandInBackend())
        X = torch.from_numpy(d["X"]).double()
        degE, degV, W = (torch.from_numpy(d[k]).double().reshape(-1) for k in ("degE", "degV", "W"))
        sl = slice(info.v_start, info.v_end)
        Y = agg.forward(X[sl], s1=degE, s2=W, a_out=degV[sl])
        G = agg.forward(X[sl], s1=degE, s2=W, a_in=degV[sl])          # transpose-backward form
        U = agg.forward(X[sl])                                          # un-scaled
        Y7 = agg.forward(X[sl][:, :7].contiguous(), s1=degE, s2=W, a_out=degV[sl])   # F % 4 != 0: the CSR-kernel path
        Y1 = PartitionedAggregator(info, TorchStandInBackend(), split_stage_a=False).forward(X[sl], s1=degE, s2=W, a_out=degV[sl])
        torch.save(dict(Y=Y, G=G, U=U, Y7=Y7, Y1=Y1, v0=info.v_start, v1=info.v_end, nb=info.num_boundary_total,
                        nint=int(info.int_edges.numel()), bytes=agg.bytes_exchanged),
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("g", ["mini", "mini_rep3"])
def test_partitioned_aggregation_matches_oracle(world, g, tmp_path):
    graph = os.path.join(os.path.dirname(__file__), "golden", f"graph_{g}.npz")
    mp.spawn(_worker, args=(world, _free_port(), graph, str(tmp_path)), nprocs=world, join=True)
    d = load_golden("graph_" + g)
    N = int(d["num_nodes"])
    parts = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert [p["v0"] for p in parts] == vertex_blocks(N, world)[:-1]
    kw = dict(s1=d["degE"], s2=d["W"])
    want = {"Y": orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], d["X"], a_out=d["degV"], **kw),
            "G": orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], d["X"], a_in=d["degV"], **kw),
            "U": orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], d["X"]),
            "Y1": orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], d["X"], a_out=d["degV"], **kw),   # one stage A over all local hyperedges
            "Y7": orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], np.ascontiguousarray(d["X"][:, :7]), a_out=d["degV"], **kw)}
    for k, w in want.items():
        got = torch.cat([p[k] for p in parts]).numpy()
        assert got.shape == w.shape and orc.rel_err(got, w) < 1e-6, k
    # every hyperedge is interior to exactly one rank or boundary; boundary rows did travel
    assert sum(p["nint"] for p in parts) + parts[0]["nb"] == int((np.diff(d["H_T_csrptr"]) > 0).sum())
    if g == "mini_rep3" and world == 3:
        # three block-diagonal replicas over three equal blocks: nothing crosses a rank boundary
        assert parts[0]["nb"] == 0 and all(p["bytes"] == 0 for p in parts)
    else:
        assert parts[0]["nb"] > 0 and all(p["bytes"] > 0 for p in parts)


def test_partition_lists_are_consistent():
    """Pure index arithmetic (no process group): send/recv lists of every rank pair mirror each other."""
    d = load_golden("graph_mini_rep3")
    N, M, P = int(d["num_nodes"]), int(d["num_edges"]), 4
    ptr, col = torch.from_numpy(d["H_T_csrptr"]), torch.from_numpy(d["H_T_colind"])
    infos = [build_partition(ptr, col, N, M, P, r) for r in range(P)]
    for r, a in enumerate(infos):
        assert int(a.int_ptr[-1]) + int(a.bnd_ptr[-1]) == int(((col >= a.v_start) & (col < a.v_end)).sum())
        for q, b in enumerate(infos):
            if q == r:
                assert a.send_rows[q].numel() == 0 and a.recv_own_pos[q].numel() == 0
                continue
            sent = a.bnd_edges[a.send_rows[q]]                           # r -> q, global ids
            got = b.bnd_edges[b.own_rows][b.recv_own_pos[r]]            # what q expects from r
            assert torch.equal(sent, got)
    # a single rank has no boundary at all
    one = build_partition(ptr, col, N, M, 1, 0)
    assert one.bnd_edges.numel() == 0 and one.int_edges.numel() == int((np.diff(d["H_T_csrptr"]) > 0).sum())
