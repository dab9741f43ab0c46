"""Host-side logic of the product (no GPU): native balancer and CSR builder are bit-exact
with the reference's vectors, and HyperGraph(device='cpu') reproduces the scipy route."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import balancer_cases, load_golden
from hypergef_b200 import HyperGraph, balance_schedule
from hypergef_b200.hypergraph import build_csr
from oracle import oracle as orc


def test_native_host_balancer_matches_reference_vectors():
    d, names = balancer_cases()
    for name in names:
        bs = balance_schedule(int(d[f"{name}__ngs"]), torch.from_numpy(d[f"{name}__csrptr"]))
        for attr, k in (("balan_key", "key"), ("balan_row", "row"), ("group_st", "st"), ("group_ed", "ed")):
            assert np.array_equal(getattr(bs, attr), d[f"{name}__{k}"]), f"{name}:{k}"
            assert getattr(bs, attr).dtype == np.int32


def test_native_host_balancer_matches_oracle_on_random_rows():
    rng = np.random.default_rng(11)
    for it in range(30):
        nrow = int(rng.integers(1, 400))
        deg = rng.integers(0, 30, size=nrow)
        if it % 3 == 0:
            deg[rng.integers(0, nrow)] = int(rng.integers(500, 5000))
        if deg.sum() == 0:
            deg[-1] = 2
        ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        ngs = int(rng.integers(1, 80))
        a, b = balance_schedule(ngs, torch.from_numpy(ptr)), orc.c_balancer(ngs, ptr)
        for k in ("balan_key", "balan_row", "group_st", "group_ed"):
            assert np.array_equal(getattr(a, k), getattr(b, k))


def test_balancer_errors_follow_the_reference():
    with pytest.raises(IndexError):                       # balancer.py:32 on an empty key list
        balance_schedule(3, torch.tensor([0, 0, 0], dtype=torch.int32))
    with pytest.raises(ValueError):
        balance_schedule(0, torch.tensor([0, 2], dtype=torch.int32))
    with pytest.raises(TypeError):
        balance_schedule(3, torch.tensor([0.0, 2.0]))
    # int64 offsets and python lists are accepted (the reference takes the scipy indptr tensor)
    assert list(balance_schedule(3, [0, 6]).balan_key) == [0, 3, 6]
    assert list(balance_schedule(3, torch.tensor([0, 4, 4])).balan_key) == [0, 3, 4]


@pytest.mark.parametrize("g", ["cora", "mini", "mini_rep3"])
def test_hypergraph_cpu_is_bit_exact_with_scipy_route(g):
    d = load_golden("graph_" + g)
    N = int(d["num_nodes"])
    data = SimpleNamespace(x=torch.zeros(N, 1), edge_index=torch.from_numpy(d["edge_index"]))
    hg = HyperGraph(data, "cpu", "synthetic", ngs=int(d["ngs"]))
    assert (hg.num_nodes, hg.num_edges, hg.nnz) == (N, int(d["num_edges"]), int(d["nnz"]))
    for attr in ("H_csrptr", "H_colind", "H_data", "H_T_csrptr", "H_T_colind", "H_T_data", "degV", "degE",
                 "group_key", "group_row", "group_start", "group_end"):
        got = getattr(hg, attr).numpy()
        assert got.dtype == d[attr].dtype and np.array_equal(got, d[attr]), attr
    assert hg.degV.shape == (N, 1) and hg.degE.shape == (hg.num_edges, 1)


def test_csr_builder_unsorted_input_duplicates_and_empty_rows():
    rng = np.random.default_rng(3)
    N, M, n = 50, 20, 400
    V, E = rng.integers(0, N, n), rng.integers(0, M, n)
    V[V == 7] = 8                      # empty row
    E[E == 3] = 4                      # empty column
    got = build_csr(torch.from_numpy(V), torch.from_numpy(E), N, M, torch.device("cpu"))
    H, H_T = orc.scipy_incidence(V, E, N, M)
    for a, b in zip(got, (H.indptr, H.indices, H.data, H_T.indptr, H_T.indices, H_T.data)):
        assert np.array_equal(a.numpy(), b.astype(a.numpy().dtype))
    with pytest.raises(ValueError):
        build_csr(torch.tensor([N]), torch.tensor([0]), N, M, torch.device("cpu"))


def test_hypergraph_argument_errors():
    data = SimpleNamespace(x=torch.zeros(4, 1), edge_index=torch.tensor([[0, 1], [1, 0]]))
    with pytest.raises(ValueError, match="num_nodes"):
        HyperGraph(data, "cpu", "cora")
    d = load_golden("graph_mini")
    data = SimpleNamespace(x=torch.zeros(int(d["num_nodes"]), 1), edge_index=torch.from_numpy(d["edge_index"]))
    with pytest.raises(KeyError):
        HyperGraph(data, "cpu", "not-a-dataset")
    assert HyperGraph(data, "cpu", "cora").ngs == 210          # hypergraph.py:74 table


def test_degree_powers_vs_torch_pow():
    """hypergraph.py:40-41 runs torch.pow(d,-0.5) / pow(d,-1) on the host.  pow(-1) is the exact
    reciprocal; pow(-0.5) is torch's vectorised rsqrt, which is NOT correctly rounded (it differs
    from float(1/sqrt(double d)) for ~25% of integer degrees, CPU-ISA dependent), so degV is an
    fp32 quantity under the 1e-5 tolerance, not a bit-exact one.  The device kernel uses the
    correctly rounded 1/sqrtf(d): within 2 ulp of what the reference computed."""
    deg = torch.arange(1, 200001, dtype=torch.float32)
    assert torch.equal(deg.pow(-1), 1.0 / deg)
    mine = 1.0 / torch.sqrt(deg)
    rel = ((deg.pow(-0.5) - mine).abs() / mine).max().item()
    assert rel <= 2.5e-7          # at most 2 ulp


def test_projection_order_cost_model():
    """ops.projection_order: where a layer's projection goes, by a byte / flop count."""
    from hypergef_b200 import ops
    N, M = 1261888, 509632                                   # the bench graph: E < N
    assert ops.projection_order(N, M, 256, 256) == "edge"    # fewer rows to project between the stages
    assert ops.projection_order(2708, 1579, 1433, 32) == "vertex"   # shrink the features first (the reference's order)
    assert ops.projection_order(N, M, 32, 256) == "after"    # aggregate the narrow features, project last
    assert ops.projection_order(N, M, 256, 7) == "vertex"    # the stage entry points need multiples of 4: never 'edge'
    assert ops.projection_order(N, M, 30, 64) in ("vertex", "after")
    # more hyperedges than vertices: projecting the hyperedge rows is never the cheapest
    assert ops.projection_order(1000, 5000, 128, 128) != "edge"


def test_data_from_members_validates():
    import pytest
    from hypergef_b200 import io as hio
    with pytest.raises(ValueError):
        hio.data_from_members([0, 1], [0], 3, None, None)
    with pytest.raises(ValueError):
        hio.data_from_members([0, 5], [0, 0], 3, None, None)
    d = hio.data_from_members([], [], 3, None, None)
    assert d.edge_index.shape == (2, 0) and d.num_hyperedges == 0

