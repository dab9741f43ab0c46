"""Parity of the CUDA path (through the C-ABI) with the oracle -- runs on the B200 box.

Bit-exact for integer work (balancer, CSR); <= 1e-5 relative (max-abs error over max-abs value,
the north star's tolerance for atomic-reordered fp32 sums) against the fp64 oracle for features.
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import balancer_cases, load_golden
import hypergef_b200 as hgef
from hypergef_b200 import HyperGraph, balance_schedule, ops, synth, _native
from hypergef_b200.hypergraph import build_csr
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(params=["auto", "two_pass", "stream"])
def kernel_form(request):
    """Run a test with the library's own choice of kernel form, then with each product form forced (the
    stream form is only chosen on its own when Y exceeds the L2).  The experimental single-launch forms live
    in the lab library and are tested by tests/lab/ (tests/test_lab.py)."""
    ops.DEFAULT_FLAGS = {"auto": 0, "two_pass": _native.HG_TWO_PASS, "stream": _native.HG_FORCE_STREAM}[request.param]
    yield request.param
    ops.DEFAULT_FLAGS = 0


def _graph(g, dev):
    d = load_golden("graph_" + g)
    data = SimpleNamespace(x=torch.zeros(int(d["num_nodes"]), 1), edge_index=torch.from_numpy(d["edge_index"]))
    return d, HyperGraph(data, dev, "synthetic", ngs=int(d["ngs"]))


def _np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ integer work: bit-exact
def test_device_balancer_bit_exact(cuda_device):
    d, names = balancer_cases()
    for name in names:
        bs = balance_schedule(int(d[f"{name}__ngs"]), torch.from_numpy(d[f"{name}__csrptr"]).to(cuda_device))
        for attr, k in (("balan_key", "key"), ("balan_row", "row"), ("group_st", "st"), ("group_ed", "ed")):
            got = getattr(bs, attr)
            assert got.is_cuda and got.dtype == torch.int32
            assert np.array_equal(_np(got), d[f"{name}__{k}"]), f"{name}:{k}"
    with pytest.raises(IndexError):
        balance_schedule(3, torch.zeros(4, dtype=torch.int32, device=cuda_device))


def test_device_balancer_equals_host_at_scale(cuda_device):
    rng = np.random.default_rng(2)
    deg = rng.geometric(0.25, size=300_000)
    deg[rng.integers(0, deg.size, 50)] = rng.integers(1000, 30000, 50)
    deg[rng.integers(0, deg.size, 1000)] = 0
    ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]).astype(np.int32))
    for ngs in (7, 210):
        h, g = balance_schedule(ngs, ptr), balance_schedule(ngs, ptr.to(cuda_device))
        for k in ("balan_key", "balan_row", "group_st", "group_ed"):
            assert np.array_equal(getattr(h, k), _np(getattr(g, k))), (ngs, k)


@pytest.mark.parametrize("g", ["cora", "mini", "mini_rep3"])
def test_device_hypergraph_bit_exact(g, cuda_device):
    d, hg = _graph(g, cuda_device)
    for attr in ("H_csrptr", "H_colind", "H_data", "H_T_csrptr", "H_T_colind", "H_T_data",
                 "group_key", "group_row", "group_start", "group_end"):
        got = _np(getattr(hg, attr))
        assert got.dtype == d[attr].dtype and np.array_equal(got, d[attr]), attr
    assert np.array_equal(_np(hg.degE), d["degE"])                        # exact reciprocal
    assert np.abs(_np(hg.degV) - d["degV"]).max() <= 2.5e-7 * d["degV"].max()   # torch rsqrt is not exact


def test_device_csr_random_unsorted_duplicates(cuda_device):
    rng = np.random.default_rng(9)
    N, M, n = 5000, 1200, 60000
    V, E = rng.integers(0, N, n), rng.integers(0, M, n)
    V[:500], E[:500] = V[500:1000], E[500:1000]            # duplicates
    E[E == 17] = 18
    got = build_csr(torch.from_numpy(V), torch.from_numpy(E), N, M, cuda_device)
    H, H_T = orc.scipy_incidence(V, E, N, M)
    for a, b in zip(got, (H.indptr, H.indices, H.data, H_T.indptr, H_T.indices, H_T.data)):
        assert np.array_equal(_np(a), b.astype(_np(a).dtype))
    host = build_csr(torch.from_numpy(V), torch.from_numpy(E), N, M, torch.device("cpu"))
    for a, b in zip(got, host):
        assert torch.equal(a.cpu(), b)
    with pytest.raises(ValueError):
        build_csr(torch.tensor([N]), torch.tensor([0]), N, M, cuda_device)


# ------------------------------------------------------------------ features: <= 1e-5 relative
@pytest.mark.parametrize("g", ["mini", "mini_rep3"])
def test_forward_matches_golden(g, cuda_device, kernel_form):
    d, hg = _graph(g, cuda_device)
    X = torch.from_numpy(d["X"]).to(cuda_device)
    W = torch.from_numpy(d["W"]).to(cuda_device)
    Y = hgef.HGNNAggr(hg, X, hg.degE, hg.degV, W, "sum")
    assert orc.rel_err(_np(Y), d["Y_hgnn_f64"]) < TOL
    Yu = hgef.UniGNNConv(hg, X)
    assert orc.rel_err(_np(Yu), d["Y_unscaled_ref_host"]) < TOL     # the reference's own host golden
    Yd = hgef.UniGNNConvdeg(hg, X, hg.degE, hg.degV)
    want = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], d["X"], s1=d["degE"], a_out=d["degV"])
    assert orc.rel_err(_np(Yd), want) < TOL
    # the reference test's 5-argument call (test/hgnn_test.py:89)
    assert hgef.HGNNAggr(hg, X, hg.degE, hg.degV, W).shape == Y.shape


@pytest.mark.parametrize("F", [1, 2, 7, 12, 32, 64, 100, 128, 200, 256, 512, 516])
def test_forward_feature_lengths(F, cuda_device, kernel_form):
    """Every vector layout (lanes/row 1..32, 1/2/4 vectors per lane) and the scalar tail path."""
    d, hg = _graph("mini", cuda_device)           # ngs=6, max hyperedge 75: light and heavy hyperedges
    N, M = hg.num_nodes, hg.num_edges
    X = torch.randn(N, F, generator=torch.Generator().manual_seed(F))
    W = 0.5 + torch.rand(M, generator=torch.Generator().manual_seed(1))
    want = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X, s1=d["degE"], s2=W, a_out=d["degV"])
    Y = ops.hgnnaggr(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_csrptr, hg.H_T_colind,
                     X.to(cuda_device), hg.degE, hg.degV, W.to(cuda_device))
    assert Y.shape == (N, F) and orc.rel_err(_np(Y), want) < TOL
    # the output buffer is NOT pre-zeroed by the caller: garbage in it must not leak through
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    out = torch.full((N, F), float("nan"), device=cuda_device)
    ops.aggregate(plan, X.to(cuda_device), s1=hg.degE, s2=W.to(cuda_device), a_out=hg.degV, out=out)
    plan.check()
    assert orc.rel_err(_np(out), want) < TOL


def test_plan_sees_heavy_hyperedges_and_schedules_agree(cuda_device):
    d, hg = _graph("mini_rep3", cuda_device)
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind,
                        hg.num_nodes, hg.num_edges)
    assert plan.canonical and plan.nseg == d["group_key"].size - 1
    w = np.ceil(np.diff(d["H_T_csrptr"]) / int(d["ngs"])).astype(int)
    assert plan.nheavy_edges == int((w > 1).sum()) and plan.nheavy_segs == int(w[w > 1].sum())
    X = torch.from_numpy(d["X"]).to(cuda_device)
    Y1 = ops.aggregate(plan, X, s1=hg.degE, a_out=hg.degV)
    # the literal reference schedule (one work unit per balancer group) through the C-ABI
    Y2 = torch.empty_like(Y1)
    _native.call("hg_aggr_groups", hg.num_nodes, hg.group_row.numel(), hg.group_key.data_ptr(),
                 hg.group_row.data_ptr(), hg.group_start.data_ptr(), hg.group_end.data_ptr(),
                 hg.H_T_colind.data_ptr(), X.data_ptr(), hg.degE.data_ptr(), None, hg.degV.data_ptr(), None,
                 Y2.data_ptr(), X.shape[1], 0, 0, torch.cuda.current_stream().cuda_stream)
    want = orc.c_aggr_groups(d["group_key"], d["group_row"], d["group_start"], d["group_end"], d["H_T_colind"],
                             d["X"], s1=d["degE"], a_out=d["degV"])
    assert orc.rel_err(_np(Y1), want) < TOL and orc.rel_err(_np(Y2), want) < TOL
    for flag in (_native.HG_FORCE_SCALAR, _native.HG_TWO_PASS, _native.HG_FORCE_STREAM):
        Y3 = ops.aggregate(plan, X, s1=hg.degE, a_out=hg.degV, flags=flag)
        assert orc.rel_err(_np(Y3), want) < TOL
    plan.check()
    # accumulate flag: Y += op(X)
    Y4 = ops.aggregate(plan, X, s1=hg.degE, a_out=hg.degV, out=Y1.clone(), flags=_native.HG_ACCUMULATE)
    assert orc.rel_err(_np(Y4), 2 * want) < TOL


def test_non_canonical_groups_keep_reference_meaning(cuda_device):
    """User-built group arrays that are not the balancer's full cross product run with the
    literal semantics of hgnnaggr_cuda.cu:14-47 (gather seg st[g], scatter seg ed[g])."""
    d, hg = _graph("mini", cuda_device)
    rng = np.random.default_rng(0)
    keep = np.sort(rng.choice(d["group_row"].size, size=d["group_row"].size // 2, replace=False))
    row, st, ed = (torch.from_numpy(d[k][keep]).to(cuda_device) for k in ("group_row", "group_start", "group_end"))
    plan = ops.get_plan(hg.group_key, row, st, ed, hg.H_T_colind, hg.num_nodes, hg.num_edges)
    assert not plan.canonical
    X = torch.from_numpy(d["X"]).to(cuda_device)
    Y = ops.aggregate(plan, X, s1=hg.degE, a_out=hg.degV)
    want = orc.c_aggr_groups(d["group_key"], d["group_row"][keep], d["group_start"][keep], d["group_end"][keep],
                             d["H_T_colind"], d["X"], s1=d["degE"], a_out=d["degV"])
    assert orc.rel_err(_np(Y), want) < TOL


def test_backward_transpose_and_reference_modes(cuda_device, kernel_form):
    d, hg = _graph("mini", cuda_device)
    N, M = hg.num_nodes, hg.num_edges
    F = 16
    gen = torch.Generator().manual_seed(5)
    X0, G0 = torch.randn(N, F, generator=gen), torch.randn(N, F, generator=gen)
    W0 = 0.5 + torch.rand(M, generator=gen)
    # fp64 autograd through the PyG-equivalent conv (model/pygnn/hgnn.py:30-37)
    E = torch.from_numpy(np.repeat(np.arange(M), np.diff(d["H_T_csrptr"]))).long()
    V = torch.from_numpy(d["H_T_colind"].astype(np.int64))
    Xd, Wd = X0.double().requires_grad_(True), W0.double().requires_grad_(True)
    Yd = orc.torch_hgnn_conv(Xd, V, E, torch.from_numpy(d["degE"]).double(), torch.from_numpy(d["degV"]).double(),
                             Wd, N, M)
    Yd.backward(G0.double())
    X = X0.to(cuda_device).requires_grad_(True)
    W = W0.to(cuda_device).requires_grad_(True)
    assert hgef.get_backward_mode() == "transpose"
    Y = hgef.HGNNAggr(hg, X, hg.degE, hg.degV, W, "sum")
    Y.backward(G0.to(cuda_device))
    assert orc.rel_err(_np(Y), Yd.detach().numpy()) < TOL
    assert orc.rel_err(_np(X.grad), Xd.grad.numpy()) < TOL
    assert orc.rel_err(_np(W.grad), Wd.grad.numpy()) < TOL           # the reference returns no W grad
    # reference mode: backward == forward applied to grad_out (hgnnaggr.cc:58-60)
    hgef.set_backward_mode("reference")
    try:
        X2 = X0.to(cuda_device).requires_grad_(True)
        hgef.HGNNAggr(hg, X2, hg.degE, hg.degV, W.detach(), "sum").backward(G0.to(cuda_device))
        want = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], G0, s1=d["degE"], s2=W0, a_out=d["degV"])
        assert orc.rel_err(_np(X2.grad), want) < TOL
    finally:
        hgef.set_backward_mode("transpose")
    # un-scaled operator is symmetric: both modes coincide (unignnaggr.cc:72-74)
    X3 = X0.to(cuda_device).requires_grad_(True)
    hgef.UniGNNConv(hg, X3).backward(G0.to(cuda_device))
    assert orc.rel_err(_np(X3.grad), orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], G0)) < TOL


def test_mean_and_max_variants(cuda_device):
    d, hg = _graph("mini", cuda_device)
    N, M, F = hg.num_nodes, hg.num_edges, 40
    gen = torch.Generator().manual_seed(8)
    X0, G0 = torch.randn(N, F, generator=gen), torch.randn(N, F, generator=gen)
    W0 = 0.5 + torch.rand(M, generator=gen)
    X, W = X0.to(cuda_device).requires_grad_(True), W0.to(cuda_device)
    Ym = hgef.HGNNAggr(hg, X, hg.degE, hg.degV, W, "mean")
    want = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X0, s1=d["degE"], s2=W0, a_out=d["degV"], reduce="mean")
    assert orc.rel_err(_np(Ym), want) < TOL
    Ym.backward(G0.to(cuda_device))
    wantg = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], G0, s1=d["degE"], s2=W0, a_out=d["degV"], reduce="mean")
    assert orc.rel_err(_np(X.grad), wantg) < TOL
    X.grad = None
    out, rec = hgef.hgnnaggr.hgnnaggr_max(hg.H_T_csrptr, hg.H_T_colind, X, hg.degE, hg.degV, W)
    wy, wrec = orc.c_aggr_max(d["H_T_csrptr"], d["H_T_colind"], X0, s1=d["degE"], s2=W0, a_out=d["degV"])
    assert orc.rel_err(_np(out), wy) < TOL and np.array_equal(_np(rec), wrec)
    out.backward(G0.to(cuda_device))
    wdx = orc.c_aggr_max_bwd(d["H_T_csrptr"], d["H_T_colind"], G0, wrec, s1=d["degE"], s2=W0, a_out=d["degV"])
    assert orc.rel_err(_np(X.grad), wdx) < TOL


@pytest.mark.parametrize("shape,F", [("cora", 32), ("pubmed", 64), ("dblp", 128), ("walmart", 32)])
def test_baseline_shapes_against_oracle(shape, F, cuda_device, kernel_form):
    """BASELINE.json configs at their literal sizes (C1-C4); walmart plants a 12345-member hyperedge."""
    data = synth.make_shape(shape, seed=0)
    hg = HyperGraph(data, cuda_device, data.dataset)
    if shape == "walmart":
        assert int((hg.H_T_csrptr[1:] - hg.H_T_csrptr[:-1]).max()) > 10_000
    X = torch.randn(hg.num_nodes, F, generator=torch.Generator().manual_seed(1))
    Y = hgef.HGNNAggr(hg, X.to(cuda_device), hg.degE, hg.degV, torch.ones(hg.num_edges, device=cuda_device))
    want = orc.c_aggr_formula(_np(hg.H_T_csrptr), _np(hg.H_T_colind), X, s1=_np(hg.degE), a_out=_np(hg.degV))
    assert orc.rel_err(_np(Y), want) < TOL
    # balancer at the literal size is bit-exact with the C oracle
    b = orc.c_balancer(hg.ngs, _np(hg.H_T_csrptr))
    assert np.array_equal(_np(hg.group_key), b.balan_key) and np.array_equal(_np(hg.group_end), b.group_ed)


def test_full_size_properties(cuda_device, kernel_form):
    """Pubmed-shaped x64 (the bench workload) at F=128: properties that need no CPU oracle --
    linearity, the adjoint identity <A x, z> = <x, A^T z>, and the column checksum
    1^T (H H^T X) = sum_e |e| (H^T X)_e."""
    data = synth.make_shape("pubmed", replicas=64, seed=0, device=cuda_device)
    hg = HyperGraph(data, cuda_device, "pubmed")
    N, F = hg.num_nodes, 128
    gen = torch.Generator(device=cuda_device).manual_seed(0)
    X, Z = (torch.randn(N, F, device=cuda_device, generator=gen) for _ in range(2))
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, hg.num_edges)
    A = lambda x, **kw: ops.aggregate(plan, x, s1=hg.degE, **kw)
    y1, y2 = A(X, a_out=hg.degV), A(Z, a_out=hg.degV)
    y3 = A(2.0 * X - 3.0 * Z, a_out=hg.degV)
    assert ((y3 - (2.0 * y1 - 3.0 * y2)).abs().max() / y3.abs().max()).item() < TOL
    lhs = (y1.double() * Z.double()).sum()
    rhs = (X.double() * A(Z, a_in=hg.degV).double()).sum()
    assert abs((lhs - rhs) / lhs).item() < 1e-6
    plan.check()
    yu = ops.aggregate(plan, X)
    deg = (hg.H_T_csrptr[1:] - hg.H_T_csrptr[:-1]).double()
    rows = torch.repeat_interleave(torch.arange(hg.num_edges, device=cuda_device), deg.long())
    xe = torch.zeros(hg.num_edges, F, dtype=torch.float64, device=cuda_device).index_add_(
        0, rows, X.double()[hg.H_T_colind.long()])
    want = (xe * deg[:, None]).sum(0)
    assert ((yu.double().sum(0) - want).abs().max() / want.abs().max()).item() < 1e-6


ST_KNOBS = ("st_slab", "st_sw", "st_l", "st_ctas", "st_occ", "st_pipe", "st_cs", "st_pdl", "stream_min_mb")


@pytest.mark.parametrize("shape,replicas", [("pubmed", 3), ("walmart", 1), ("dblp", 2)])
def test_stream_form_configurations(shape, replicas, cuda_device):
    """The stream form (the shipped large-graph form: stage A and stage B as two launches, B a programmatic
    dependent launch of A, self-resetting ticket counters) under every geometry knob -- item length, column
    slabs, sub-warp width, occupancy, pipelining, store hint, PDL on / off -- forward and transposed, EVERY
    feature length against the fp64 C oracle, element-wise against the sum of the magnitudes of the terms as
    well as against the largest value.  Bit-identical run to run (fixed summation order, no atomics for light
    hyperedges) where the graph has no heavy hyperedge."""
    data = synth.make_shape(shape, replicas=replicas, seed=3)
    hg = HyperGraph(data, cuda_device, data.dataset)
    N, M = hg.num_nodes, hg.num_edges
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    W = torch.rand(M, device=cuda_device) + 0.5
    ptr, ind = _np(hg.H_T_csrptr), _np(hg.H_T_colind)
    s_edge = _np(hg.degE).ravel() * _np(W)
    combos = [dict(), dict(st_pdl=0), dict(st_l=16), dict(st_l=48, st_slab=32), dict(st_l=256, st_slab=128, st_ctas=1),
              dict(st_occ=4), dict(st_pipe=0, st_l=32), dict(st_pipe=1), dict(st_sw=16), dict(st_sw=8, st_occ=4),
              dict(st_sw=4, st_cs=0), dict(st_slab=64, st_l=16, st_pdl=0)]
    try:
        for F in (4, 20, 32, 64, 100, 128, 256, 384, 512, 640):
            X = torch.randn(N, F, device=cuda_device)
            want = orc.c_aggr_formula(ptr, ind, X.cpu(), s1=s_edge, a_out=_np(hg.degV))
            bound = orc.c_aggr_formula(ptr, ind, X.cpu().abs(), s1=np.abs(s_edge), a_out=np.abs(_np(hg.degV)))
            want_t = orc.c_aggr_formula(ptr, ind, X.cpu(), s1=s_edge, a_in=_np(hg.degV))
            for knobs in combos:
                ops.tune(**{k: None for k in ST_KNOBS})
                ops.tune(**knobs)
                out = torch.full((N, F), float("nan"), device=cuda_device)
                ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, out=out, flags=_native.HG_FORCE_STREAM)
                assert orc.rel_err(_np(out), want) < TOL, (shape, F, knobs)
                assert orc.rel_err_terms(_np(out), want, bound) < TOL, (shape, F, knobs, "element-wise")
                out_t = ops.aggregate(plan, X, s1=hg.degE, s2=W, a_in=hg.degV, flags=_native.HG_FORCE_STREAM)
                assert orc.rel_err(_np(out_t), want_t) < TOL, (shape, F, knobs, "transposed")
                if plan.nheavy_edges == 0:
                    again = ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, flags=_native.HG_FORCE_STREAM)
                    assert torch.equal(again, out), (shape, F, knobs, "run-to-run")
                plan.check()
    finally:
        ops.tune(**{k: None for k in ST_KNOBS})


def test_odd_feature_lengths_on_a_large_graph_use_padded_rows(cuda_device):
    """F % 4 != 0 on a graph whose Y exceeds the L2 (threshold lowered through the tuning table so that a
    small graph takes that route): the stream kernels run on rows padded to the next multiple of 4 instead of
    the scalar-atomic path; the result equals the oracle and NaNs in the output buffer do not leak."""
    data = synth.make_shape("dblp", replicas=2, seed=4)
    hg = HyperGraph(data, cuda_device, data.dataset)
    N, M = hg.num_nodes, hg.num_edges
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    ptr, ind = _np(hg.H_T_csrptr), _np(hg.H_T_colind)
    try:
        ops.tune(stream_min_mb=0)
        for F in (1, 3, 7, 30, 101):
            X = torch.randn(N, F, device=cuda_device)
            want = orc.c_aggr_formula(ptr, ind, X.cpu(), s1=_np(hg.degE), a_out=_np(hg.degV))
            k0 = plan.kernels_launched()
            out = torch.full((N, F), float("nan"), device=cuda_device)
            ops.aggregate(plan, X, s1=hg.degE, a_out=hg.degV, out=out)
            plan.check()
            assert plan.kernels_launched() - k0 == 4 + (1 if plan.nheavy_segs else 0)   # pad, A, B, unpad
            assert orc.rel_err(_np(out), want) < TOL, F
    finally:
        ops.tune(stream_min_mb=None)


def test_plan_reserve_and_capture_never_allocate(cuda_device):
    """hg_plan_reserve sizes the per-call buffers once; a call that would have to grow one while the stream is
    being captured is refused (ValueError) instead of allocating inside the capture; a buffer that grows is not
    freed under a graph that still uses it (the earlier graph replays correctly afterwards)."""
    data = synth.make_shape("pubmed", replicas=2, seed=6)
    hg = HyperGraph(data, cuda_device, data.dataset)
    N, M = hg.num_nodes, hg.num_edges
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    flags = _native.HG_FORCE_STREAM
    X32 = torch.randn(N, 32, device=cuda_device)
    Y32 = torch.empty_like(X32)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        plan.reserve(32)
        g32 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g32, stream=side):      # first use of the plan, captured: reserve made it legal
            ops.aggregate(plan, X32, s1=hg.degE, a_out=hg.degV, out=Y32, flags=flags)
        X64 = torch.randn(N, 64, device=cuda_device)
        Y64 = torch.empty_like(X64)
        g64 = torch.cuda.CUDAGraph()
        with pytest.raises(ValueError):
            with torch.cuda.graph(g64, stream=side):  # wider F without a reserve: must refuse, not allocate
                ops.aggregate(plan, X64, s1=hg.degE, a_out=hg.degV, out=Y64, flags=flags)
    torch.cuda.synchronize()
    plan.reserve(64)                                  # grows the buffer g32 was captured with: it must stay alive
    want64 = ops.aggregate(plan, X64, s1=hg.degE, a_out=hg.degV, flags=_native.HG_TWO_PASS)
    got64 = ops.aggregate(plan, X64, s1=hg.degE, a_out=hg.degV, flags=flags)
    assert ((got64 - want64).abs().max() / want64.abs().max()).item() < TOL
    Y32.fill_(float("nan"))
    g32.replay()
    torch.cuda.synchronize()
    want32 = ops.aggregate(plan, X32, s1=hg.degE, a_out=hg.degV, flags=_native.HG_TWO_PASS)
    assert ((Y32 - want32).abs().max() / want32.abs().max()).item() < TOL
    plan.check()


@pytest.mark.parametrize("form", ["two_pass", "stream"])
def test_cuda_graph_capture_and_replay(form, cuda_device):
    """Every kernel form is capture-safe after one eager warm-up call (the first call of a plan may allocate its
    scratch): a captured aggregation replays correctly on new contents of the same input buffer."""
    flags = {"two_pass": _native.HG_TWO_PASS, "stream": _native.HG_FORCE_STREAM}[form]
    data = synth.make_shape("pubmed", replicas=2, seed=5)
    hg = HyperGraph(data, cuda_device, data.dataset)
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, hg.num_nodes,
                        hg.num_edges)
    F = 64
    X = torch.randn(hg.num_nodes, F, device=cuda_device)
    Y = torch.empty_like(X)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ops.aggregate(plan, X, s1=hg.degE, a_out=hg.degV, out=Y, flags=flags)      # warm-up: allocations happen here
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            ops.aggregate(plan, X, s1=hg.degE, a_out=hg.degV, out=Y, flags=flags)
    for seed in (1, 2):
        X.copy_(torch.randn(hg.num_nodes, F, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(seed)))
        Y.fill_(float("nan"))
        graph.replay()
        torch.cuda.synchronize()
        want = ops.aggregate(plan, X, s1=hg.degE, a_out=hg.degV, flags=_native.HG_TWO_PASS)
        assert ((Y - want).abs().max() / want.abs().max()).item() < TOL, (form, seed)
    plan.check()


# ------------------------------------------------------------------ error behaviour
def test_errors_raise_instead_of_aborting(cuda_device):
    d, hg = _graph("mini", cuda_device)
    X = torch.from_numpy(d["X"]).to(cuda_device)
    with pytest.raises(TypeError):
        hgef.UniGNNConv(hg, X.double())
    with pytest.raises(ValueError):
        hgef.UniGNNConv(hg, X[:-1])
    with pytest.raises(ValueError):
        hgef.UniGNNConvdeg(hg, X, hg.degE[:-1], hg.degV)
    bad = hg.H_T_colind.clone()
    bad[3] = hg.num_nodes                                   # out-of-range vertex id
    with pytest.raises(_native.HgefGraphError):
        ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, bad, hg.num_nodes, hg.num_edges)
    badk = hg.group_key.clone()
    badk[2] = badk[1] - 1                                   # key not monotone
    with pytest.raises(_native.HgefGraphError):
        ops.get_plan(badk, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, hg.num_nodes, hg.num_edges)
    assert _native.lib().hg_device_cc(0) == 100             # B200 = sm_100


# ------------------------------------------------------------------ partitioned (multi-GPU) building blocks
@pytest.mark.parametrize("F", [7, 32, 256])
def test_stage_kernels_and_two_rank_emulation(F, cuda_device):
    """hg_edge_reduce / hg_edge_scatter against the oracle, then the whole partitioned algorithm for
    two ranks emulated on one GPU (the exchange done by hand from the PartitionInfo lists; the real
    all_to_all path is covered on CPU by tests/test_partition_cpu.py and on 2 GPUs by tools/run_partition.py)."""
    from hypergef_b200.partition import CudaBackend, build_partition
    d, hg = _graph("mini_rep3", cuda_device)
    N, M = hg.num_nodes, hg.num_edges
    gen = torch.Generator().manual_seed(F)
    X0 = torch.randn(N, F, generator=gen)
    W0 = 0.5 + torch.rand(M, generator=gen)
    X, W = X0.to(cuda_device), W0.to(cuda_device)
    be = CudaBackend(cuda_device, ngs=int(d["ngs"]))
    # stage 1 + stage 2 over the full CSR == the fused operator
    P = be.edge_reduce(hg.H_T_csrptr, hg.H_T_colind, X, None)
    rows = np.repeat(np.arange(M), np.diff(d["H_T_csrptr"]))
    wantP = np.zeros((M, F)); np.add.at(wantP, rows, X0.numpy().astype(np.float64)[d["H_T_colind"]])
    assert orc.rel_err(_np(P), wantP) < TOL
    Y = torch.zeros(N, F, device=cuda_device)
    be.edge_scatter(hg.H_T_csrptr, hg.H_T_colind, P, (hg.degE.reshape(-1) * W).contiguous(), hg.degV.reshape(-1), Y)
    want = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X0, s1=d["degE"], s2=W0, a_out=d["degV"])
    assert orc.rel_err(_np(Y), want) < TOL
    # two ranks on one device
    world = 2
    infos = [build_partition(hg.H_T_csrptr, hg.H_T_colind, N, M, world, r) for r in range(world)]
    assert infos[0].num_boundary_total > 0
    s = (hg.degE.reshape(-1) * W)
    degV = hg.degV.reshape(-1)
    Ps = [be.edge_reduce(i.bnd_ptr, i.bnd_ind, X[i.v_start:i.v_end].contiguous(), None) for i in infos]
    Qs = [p.clone() for p in Ps]
    for r, a in enumerate(infos):                 # owners add the partial rows their peers send
        for q, b in enumerate(infos):
            if q != r:
                Qs[r][a.own_rows[a.recv_own_pos[q]]] += Ps[q][b.send_rows[r]]
    for r, a in enumerate(infos):                 # and send the completed rows back
        for q, b in enumerate(infos):
            if q != r:
                Qs[r][a.send_rows[q]] = Qs[q][b.own_rows[b.recv_own_pos[r]]]
    outs = []
    for r, a in enumerate(infos):
        Xl, dl = X[a.v_start:a.v_end].contiguous(), degV[a.v_start:a.v_end].contiguous()
        plan = be.prepare_interior(a.int_ptr, a.int_ind, a.num_local, a.int_edges.numel())
        Yl = torch.full((a.num_local, F), float("nan"), device=cuda_device)
        be.interior(plan, Xl, hg.degE.reshape(-1)[a.int_edges].contiguous(), W[a.int_edges].contiguous(), dl, None, Yl)
        be.edge_scatter(a.bnd_ptr, a.bnd_ind, Qs[r], s[a.bnd_edges].contiguous(), dl, Yl)
        outs.append(Yl)
    assert orc.rel_err(_np(torch.cat(outs)), want) < TOL


def test_partitioned_aggregator_single_rank_both_paths(cuda_device):
    """PartitionedAggregator with one rank (every hyperedge interior): the balanced-stage path (F % 4 == 0:
    hg_plan_edge_reduce -> hg_plan_edge_scatter over the local plans) and the CSR-kernel path (other F) against the
    oracle; the multi-rank exchange around them is covered by tests/test_partition_cpu.py (gloo) and, on a box with
    two GPUs, by tests/test_convs.py::test_partitioned_aggregation_over_nccl_matches_single_gpu."""
    from hypergef_b200.partition import CudaBackend, PartitionedAggregator, build_partition
    d, hg = _graph("mini_rep3", cuda_device)
    N, M = hg.num_nodes, hg.num_edges
    info = build_partition(hg.H_T_csrptr, hg.H_T_colind, N, M, 1, 0)
    agg = PartitionedAggregator(info, CudaBackend(cuda_device, ngs=int(d["ngs"])))
    assert agg.plan_all is not None and agg.plan_bnd is None and info.num_boundary_total == 0
    gen = torch.Generator().manual_seed(11)
    W0 = 0.5 + torch.rand(M, generator=gen)
    for F in (8, 7, 64):
        X0 = torch.randn(N, F, generator=gen)
        Y = agg.forward(X0.to(cuda_device), s1=hg.degE, s2=W0.to(cuda_device), a_out=hg.degV)
        want = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X0, s1=d["degE"], s2=W0, a_out=d["degV"])
        assert orc.rel_err(_np(Y), want) < TOL, F
        G = agg.forward(X0.to(cuda_device), s1=hg.degE, a_in=hg.degV)
        wantg = orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X0, s1=d["degE"], a_in=d["degV"])
        assert orc.rel_err(_np(G), wantg) < TOL, F


def test_addresses_beyond_int32(cuda_device):
    """C5-shaped graph (window locality) at 1/10 scale with F=512: N*F = 2.56e9 > 2^31, the case the
    reference's int32 address arithmetic (hgnnaggr_cuda.cu:34,44) cannot represent.  Checked by exact fp64
    recomputation of sampled output rows (the full C5, N*F = 1.28e10, is tools/c5_single.py)."""
    import dataclasses
    shape = dataclasses.replace(synth.SHAPES["c5"], num_nodes=5_000_000, num_edges=1_000_000)
    data = synth.make_shape("c5", seed=0, device=cuda_device, shape=shape)
    hg = HyperGraph(data, cuda_device, "synthetic", ngs=shape.ngs)
    del data
    N, M, F = hg.num_nodes, hg.num_edges, 512
    assert N * F > 2 ** 31
    gen = torch.Generator(device=cuda_device).manual_seed(2)
    X = torch.randn(N, F, device=cuda_device, generator=gen)
    Y = hgef.HGNNAggr(hg, X, hg.degE, hg.degV, torch.ones(M, device=cuda_device))
    degE, degV = hg.degE.reshape(-1).double(), hg.degV.reshape(-1).double()
    Hp, Hc, Tp, Tc = hg.H_csrptr.long(), hg.H_colind.long(), hg.H_T_csrptr.long(), hg.H_T_colind.long()
    sample = torch.cat([torch.randint(0, N, (64,), device=cuda_device, generator=gen),
                        torch.tensor([0, N - 1, N - 2], device=cuda_device)])
    for v in sample.tolist():
        row = torch.zeros(F, dtype=torch.float64, device=cuda_device)
        for e in Hc[Hp[v]:Hp[v + 1]].tolist():
            row += degE[e] * X[Tc[Tp[e]:Tp[e + 1]]].double().sum(0)
        row *= degV[v]
        err = ((Y[v].double() - row).abs().max() / row.abs().max().clamp_min(1e-30)).item()
        assert err < TOL, (v, err)


def test_host_pipeline_matches_device_path(cuda_device):
    """The host-buffer API (pinned upload / launch / download overlapped over three streams)."""
    d, hg = _graph("mini_rep3", cuda_device)
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind,
                        hg.num_nodes, hg.num_edges)
    pipe = ops.HostPipeline(plan)           # wide matrices go up and down in column slabs of 128
    outs, wants = [], []
    for F in (8, 32, 128, 200, 300):
        X = torch.randn(hg.num_nodes, F, generator=torch.Generator().manual_seed(F)).pin_memory()
        outs.append(pipe.submit(X, None, s1=hg.degE, a_out=hg.degV))
        wants.append(orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X, s1=d["degE"], a_out=d["degV"]))
    pipe.wait()
    for o, w in zip(outs, wants):
        assert not o.is_cuda and orc.rel_err(o.numpy(), w) < TOL
    X = torch.randn(hg.num_nodes, 72)       # pageable host memory, slabs of 32 columns with a ragged tail
    whole = ops.HostPipeline(plan, col_slab=0).submit(X, None, s1=hg.degE, a_out=hg.degV)
    slabs = ops.HostPipeline(plan, col_slab=32)
    sl = slabs.submit(X, None, s1=hg.degE, a_out=hg.degV)
    slabs.wait()
    torch.cuda.synchronize()
    assert orc.rel_err(sl.numpy(), whole.numpy()) < TOL
    X = torch.randn(hg.num_nodes, 16)
    assert orc.rel_err(ops.aggregate_host(plan, X, s1=hg.degE, a_out=hg.degV).numpy(),
                       orc.c_aggr_formula(d["H_T_csrptr"], d["H_T_colind"], X, s1=d["degE"], a_out=d["degV"])) < TOL
    with pytest.raises(TypeError):
        ops.aggregate_host(plan, X.cuda())


def test_ragged_and_degenerate_graphs(cuda_device):
    """Edge cases: single-member hyperedges, isolated vertices (degV inf -> 1), a hyperedge holding every
    vertex, one-hyperedge and one-vertex graphs, ngs = 1 (every member its own segment: w = deg, w^2 groups)."""
    def build(members, N, ngs):
        V = torch.tensor([v for e, ms in enumerate(members) for v in ms])
        E = torch.tensor([e for e, ms in enumerate(members) for _ in ms])
        order = torch.argsort(V * len(members) + E)
        data = SimpleNamespace(x=torch.zeros(N, 1), edge_index=synth.incidence_to_edge_index(V[order], E[order], N))
        return HyperGraph(data, cuda_device, "synthetic", ngs=ngs)
    cases = [
        ([[0], [3], [1, 2, 3], [5]], 8, 2),                    # singletons + isolated vertices 4, 6, 7
        ([list(range(40))], 40, 7),                            # one hyperedge with every vertex, 6 segments
        ([[0]], 1, 3),                                         # 1 x 1
        ([[0, 1, 2, 3, 4], [2, 3], [4, 0]], 5, 1),             # ngs = 1
        ([list(range(i, i + 3)) for i in range(0, 60, 3)] + [list(range(0, 60, 2))], 64, 4),
    ]
    for members, N, ngs in cases:
        hg = build(members, N, ngs)
        ptr, ind = _np(hg.H_T_csrptr), _np(hg.H_T_colind)
        b = orc.c_balancer(ngs, ptr)
        assert np.array_equal(_np(hg.group_key), b.balan_key) and np.array_equal(_np(hg.group_row), b.balan_row)
        assert np.array_equal(_np(hg.group_start), b.group_st) and np.array_equal(_np(hg.group_end), b.group_ed)
        assert torch.isfinite(hg.degV).all()
        for F in (4, 5, 32):
            X = torch.randn(N, F, generator=torch.Generator().manual_seed(N + F))
            want = orc.c_aggr_formula(ptr, ind, X, s1=_np(hg.degE), a_out=_np(hg.degV))
            out = torch.full((N, F), float("nan"), device=cuda_device)
            plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, hg.num_edges)
            for flags in (0, _native.HG_TWO_PASS, _native.HG_FORCE_STREAM):
                ops.aggregate(plan, X.to(cuda_device), s1=hg.degE, a_out=hg.degV, out=out, flags=flags)
                assert orc.rel_err(_np(out), want) < TOL or np.abs(want).max() == 0, (len(members), N, ngs, F, flags)
            plan.check()
