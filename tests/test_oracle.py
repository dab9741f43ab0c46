"""The oracle is pinned: every restatement agrees with the golden vectors produced by running
the reference (oracle/make_golden.py) and, where oracle/_ref was built, with the reference's
own C++ compiled in place."""
import numpy as np
import pytest
import torch

from conftest import balancer_cases, load_golden
from oracle import oracle as orc

GRAPHS = ["cora", "mini", "mini_rep3"]


def _eq(bs, d, name):
    for attr, k in (("balan_key", "key"), ("balan_row", "row"), ("group_st", "st"), ("group_ed", "ed")):
        got = np.asarray(getattr(bs, attr), dtype=np.int32)
        assert np.array_equal(got, d[f"{name}__{k}"]), f"{name}:{k}"


def test_balancer_restatements_match_reference_vectors():
    d, names = balancer_cases()
    assert len(names) >= 14
    for name in names:
        ngs, ptr = int(d[f"{name}__ngs"]), d[f"{name}__csrptr"]
        _eq(orc.c_balancer(ngs, ptr), d, name)
        _eq(orc.py_balancer(ngs, ptr), d, name)


def test_balancer_survey_known_answers():
    # produced by running HyperGsys/balancer.py
    bs = orc.c_balancer(3, [0, 5, 5, 7, 14])
    assert list(bs.balan_key) == [0, 3, 5, 7, 10, 13, 14]
    assert list(bs.balan_row) == [0, 0, 0, 0, 2] + [3] * 9
    assert list(bs.group_st) == [0, 1, 0, 1, 2, 3, 4, 5, 3, 4, 5, 3, 4, 5]
    assert list(bs.group_ed) == [0, 0, 1, 1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5]
    assert list(orc.c_balancer(3, [0, 4, 4]).balan_key) == [0, 3, 4]
    assert list(orc.c_balancer(3, [0, 6]).balan_key) == [0, 3, 6]


def test_balancer_empty_matrix_raises_like_reference():
    with pytest.raises(IndexError):
        orc.c_balancer(3, [0, 0, 0])
    with pytest.raises(IndexError):
        orc.py_balancer(3, [0, 0, 0])


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built (no /root/reference)")
def test_balancer_matches_compiled_reference_twin():
    rng = np.random.default_rng(5)
    for _ in range(20):
        nrow = int(rng.integers(1, 200))
        deg = rng.integers(0, 50, size=nrow)
        deg[rng.integers(0, nrow)] = int(rng.integers(1, 3000))
        ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        ngs = int(rng.integers(1, 64))
        a, b = orc.c_balancer(ngs, ptr), orc.ref_balancer(ngs, ptr)
        for k in ("balan_key", "balan_row", "group_st", "group_ed"):
            assert np.array_equal(getattr(a, k), getattr(b, k))


@pytest.mark.parametrize("g", GRAPHS)
def test_csr_restatement_matches_scipy_golden(g):
    d = load_golden("graph_" + g)
    N, M = int(d["num_nodes"]), int(d["num_edges"])
    V, E, num_edges, nnz = orc.split_edge_index(torch.from_numpy(d["edge_index"]), N)
    assert num_edges == M and nnz == int(d["nnz"])
    c = orc.c_csr_from_coo(V.numpy(), E.numpy(), N, M)
    for got, key in ((c.indptr, "H_csrptr"), (c.indices, "H_colind"), (c.data, "H_data"),
                     (c.t_indptr, "H_T_csrptr"), (c.t_indices, "H_T_colind"), (c.t_data, "H_T_data")):
        assert np.array_equal(got, d[key]), key
    # and scipy itself still produces the stored arrays (pins the scipy version in use)
    H, H_T = orc.scipy_incidence(V.numpy(), E.numpy(), N, M)
    assert np.array_equal(H.indptr, d["H_csrptr"]) and np.array_equal(H_T.indices, d["H_T_colind"])
    degV, degE = orc.scipy_degrees(H)
    assert np.array_equal(degV.numpy(), d["degV"]) and np.array_equal(degE.numpy(), d["degE"])


@pytest.mark.parametrize("g", ["mini", "mini_rep3"])
def test_aggregation_restatements_match_reference_host_golden(g):
    d = load_golden("graph_" + g)
    X = d["X"]
    # un-scaled: the reference's own hyperaggr_reference_host output (compiled in place)
    want = d["Y_unscaled_ref_host"]
    got = orc.c_hyperaggr_host(d["H_csrptr"], d["H_colind"], d["H_T_csrptr"], d["H_T_colind"], X)
    assert np.array_equal(got, want)           # same summation order in fp32 -> bit-exact
    grp = orc.c_aggr_groups(d["group_key"], d["group_row"], d["group_start"], d["group_end"],
                            d["H_T_colind"], X)
    assert orc.rel_err(grp, want) < 1e-6
    frm = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X)
    assert orc.rel_err(frm, want) < 1e-6
    # scaled HGNN: group semantics == two-step formula == stored fp64 golden
    kw = dict(s1=d["degE"], s2=d["W"], a_out=d["degV"])
    grp = orc.c_aggr_groups(d["group_key"], d["group_row"], d["group_start"], d["group_end"],
                            d["H_T_colind"], X, **kw)
    assert orc.rel_err(grp, d["Y_hgnn_f64"]) < 1e-12
    assert orc.rel_err(orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X, **kw), d["Y_hgnn_f64"]) < 1e-12


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built (no /root/reference)")
def test_aggregation_matches_compiled_reference_host():
    d = load_golden("graph_mini")
    got = orc.ref_hyperaggr_host(d["H_csrptr"], d["H_colind"], d["H_T_csrptr"], d["H_T_colind"], d["X"])
    assert np.array_equal(got, d["Y_unscaled_ref_host"])
    # W gradient core (check.cuh:116-143) vs the restatement with all scales off
    G = np.random.default_rng(1).standard_normal(d["X"].shape).astype(np.float32)
    ref = orc.ref_weight_grad(d["H_T_csrptr"], d["H_T_colind"], G, d["X"])
    mine = orc.c_weight_grad(d["H_T_csrptr"], d["H_T_colind"], d["X"], G)
    assert orc.rel_err(ref, mine) < 1e-5


def test_torch_restatement_of_pyg_conv_matches_formula():
    d = load_golden("graph_mini_rep3")
    N, M = int(d["num_nodes"]), int(d["num_edges"])
    V, E, _, _ = orc.split_edge_index(torch.from_numpy(d["edge_index"]), N)
    Y = orc.torch_hgnn_conv(torch.from_numpy(d["X"]).double(), V, E, torch.from_numpy(d["degE"]).double(),
                            torch.from_numpy(d["degV"]).double(), torch.from_numpy(d["W"]).double(), N, M)
    assert orc.rel_err(Y.numpy(), d["Y_hgnn_f64"]) < 1e-12


def test_transpose_backward_is_the_autograd_gradient():
    """dX = H S H^T diag(degV) dY (a_in on the gather side), checked against
    torch autograd through the PyG-equivalent conv in fp64."""
    d = load_golden("graph_mini")
    N, M = int(d["num_nodes"]), int(d["num_edges"])
    H_T_ptr, H_T_ind = d["H_T_csrptr"], d["H_T_colind"]
    E = torch.from_numpy(np.repeat(np.arange(M), np.diff(H_T_ptr))).long()
    V = torch.from_numpy(H_T_ind.astype(np.int64))
    X = torch.from_numpy(d["X"]).double().requires_grad_(True)
    degE, degV, W = (torch.from_numpy(d[k]).double() for k in ("degE", "degV", "W"))
    Y = orc.torch_hgnn_conv(X, V, E, degE, degV, W, N, M)
    G = torch.randn(Y.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    Y.backward(G)
    mine = orc.c_aggr_formula(H_T_ptr, H_T_ind, G.float().numpy(), s1=d["degE"], s2=d["W"], a_in=d["degV"])
    assert orc.rel_err(mine, X.grad.numpy()) < 1e-6
