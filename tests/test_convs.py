"""The conv modules and the 2-layer model the epoch-ms metric is quoted on, and the
reference's own GPU kernels as a second opinion (X3) -- runs on the B200 box.

Model: ``HGsysHGNN`` (model/gnn.py:110-134) over ``HyperGsysHGNN`` (model/ugsys/hgnn.py:7-27), and the
UniGIN / UniGCNII convs (model/ugsys/unigin.py:7-26, unigcnii.py:7-26), forward and backward, against the
oracle conv of model/pygnn/hgnn.py:25-38 (``oracle.torch_hgnn_conv``) stacked with the same weights in fp64.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

import hypergef_b200 as hgef
from hypergef_b200 import HyperGraph, convs, ops, synth, _native
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _setup(dev, shape="cora", replicas=1, seed=0):
    data = synth.make_shape(shape, replicas=replicas, seed=seed)
    hg = HyperGraph(data, dev, data.dataset)
    N = hg.num_nodes
    V, E, M, Z = orc.split_edge_index(data.edge_index, N)
    return data, hg, V, E, N, M


def _oracle_layer(X64, Wt64, V, E, degE, degV, N, M, mode):
    """one conv in fp64 on the CPU: Linear (no bias) then the two-step formula"""
    XW = X64 @ Wt64.t()
    if mode == "hgnn":
        return orc.torch_hgnn_conv(XW, V, E, degE, degV, torch.ones(M, dtype=torch.float64), N, M)
    if mode == "unscaled":
        return orc.torch_hgnn_conv(XW, V, E, None, None, None, N, M)
    raise ValueError(mode)


def test_two_layer_hgnn_forward_backward_matches_stacked_oracle(cuda_device):
    """HGsysHGNN (eval mode: dropout off) forward + gradients of every weight and of the input."""
    data, hg, V, E, N, M = _setup(cuda_device, "cora")
    torch.manual_seed(1)
    nfeat, nhid, ncls = 64, 32, 7
    model = convs.HGsysHGNN(None, hg, nfeat, nhid, ncls).to(cuda_device).eval()
    X = torch.randn(N, nfeat, device=cuda_device, requires_grad=True)
    y = torch.randint(0, ncls, (N,), device=cuda_device)
    out = model(X)
    loss = Fn.nll_loss(out, y)
    loss.backward()

    degE, degV = hg.degE.double().cpu(), hg.degV.double().cpu()
    X64 = X.detach().double().cpu().requires_grad_(True)
    W1 = model.convs[0].W.weight.detach().double().cpu().requires_grad_(True)
    W2 = model.conv_out.W.weight.detach().double().cpu().requires_grad_(True)
    h = torch.relu(_oracle_layer(X64, W1, V, E, degE, degV, N, M, "hgnn"))
    want = Fn.log_softmax(_oracle_layer(h, W2, V, E, degE, degV, N, M, "hgnn"), dim=1)
    wloss = Fn.nll_loss(want, y.cpu())
    wloss.backward()
    assert orc.rel_err(out.detach().cpu().numpy(), want.detach().numpy()) < TOL
    assert abs(loss.item() - wloss.item()) < 1e-5 * max(1.0, abs(wloss.item()))
    assert orc.rel_err(X.grad.cpu().numpy(), X64.grad.numpy()) < 5e-5
    assert orc.rel_err(model.convs[0].W.weight.grad.cpu().numpy(), W1.grad.numpy()) < 5e-5
    assert orc.rel_err(model.conv_out.W.weight.grad.cpu().numpy(), W2.grad.numpy()) < 5e-5


@pytest.mark.parametrize("shape", ["cora", "pubmed", "walmart"])
def test_plan_stage_entry_points(shape, cuda_device):
    """hg_plan_edge_reduce / hg_plan_edge_scatter: each stage against the fp64 oracle, and the two chained are
    identical to the fused call's stream form (the same kernels with Xe in a caller buffer; bit for bit unless a
    hyperedge is split over several segments)."""
    data, hg, V, E, N, M = _setup(cuda_device, shape)
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    degE, degV = hg.degE.double().cpu().reshape(-1), hg.degV.double().cpu().reshape(-1)
    g = torch.Generator().manual_seed(3)
    W = (torch.rand(M, generator=g) + 0.5).to(cuda_device)
    for F in (4, 32, 100, 256):
        X0 = torch.randn(N, F, generator=g)
        X = X0.to(cuda_device)
        Xe = ops.edge_reduce(plan, X, s1=hg.degE, s2=W)
        want_e = torch.zeros(M, F, dtype=torch.float64).index_add_(0, E, X0.double()[V]) * (degE * W.double().cpu())[:, None]
        assert Xe.shape == (M, F)
        assert orc.rel_err(Xe.cpu().numpy(), want_e.numpy()) < TOL, (shape, F)
        Y = ops.edge_scatter(plan, Xe, a_out=hg.degV)
        want_y = torch.zeros(N, F, dtype=torch.float64).index_add_(0, V, want_e[E]) * degV[:, None]
        assert orc.rel_err(Y.cpu().numpy(), want_y.numpy()) < TOL, (shape, F)
        fused = ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, flags=_native.HG_FORCE_STREAM)
        if plan.nheavy_edges == 0:
            assert torch.equal(Y, fused), (shape, F)
        else:   # hyperedges split over several segments are summed with vector reductions: the order varies run to run
            assert float((Y - fused).abs().max() / fused.abs().max()) < TOL, (shape, F)
        # gather-side weight (the transposed backward's stage A) and caller-provided outputs
        out_e = torch.full((M, F), float("nan"), device=cuda_device)
        ops.edge_reduce(plan, X, a_in=hg.degV, out=out_e)
        want_t = torch.zeros(M, F, dtype=torch.float64).index_add_(0, E, (X0.double() * degV[:, None])[V])
        assert orc.rel_err(out_e.cpu().numpy(), want_t.numpy()) < TOL, (shape, F)
    plan.check()
    with pytest.raises(ValueError):
        ops.edge_reduce(plan, torch.randn(N, 6, device=cuda_device))         # not a multiple of 4
    with pytest.raises(ValueError):
        ops.edge_scatter(plan, torch.randn(M + 1, 8, device=cuda_device))    # wrong row count


@pytest.mark.parametrize("shape,replicas", [("cora", 1), ("pubmed", 4)])
def test_projected_layer_every_order_matches_the_oracle(shape, replicas, cuda_device):
    """Y = degV H (degE W) H^T (X Theta) with Theta applied to the vertex rows (the reference's
    order), to the hyperedge rows between the stages, or to the aggregated rows -- forward and the gradients of X
    and Theta against the fp64 formula."""
    data, hg, V, E, N, M = _setup(cuda_device, shape, replicas)
    degE, degV = hg.degE.double().cpu(), hg.degV.double().cpu()
    g = torch.Generator().manual_seed(5)
    Wd = torch.ones(M, device=cuda_device)
    for f_in, f_out in ((64, 32), (32, 64), (128, 128)):
        X0 = torch.randn(N, f_in, generator=g)
        T0 = torch.randn(f_in, f_out, generator=g) / f_in ** 0.5
        G0 = torch.randn(N, f_out, generator=g)
        X64, T64 = X0.double().requires_grad_(True), T0.double().requires_grad_(True)
        want = orc.torch_hgnn_conv(X64 @ T64, V, E, degE, degV, torch.ones(M, dtype=torch.float64), N, M)
        want.backward(G0.double())
        for order in ("vertex", "edge", "after", "auto"):
            X = X0.to(cuda_device).requires_grad_(True)
            T = T0.to(cuda_device).requires_grad_(True)
            Y = ops.projected_aggregate(hg, X, T, hg.degE, hg.degV, Wd, order=order)
            Y.backward(G0.to(cuda_device))
            assert orc.rel_err(Y.detach().cpu().numpy(), want.detach().numpy()) < TOL, (order, f_in, f_out)
            assert orc.rel_err(X.grad.cpu().numpy(), X64.grad.numpy()) < 5e-5, (order, f_in, f_out)
            assert orc.rel_err(T.grad.cpu().numpy(), T64.grad.numpy()) < 5e-5, (order, f_in, f_out)
    # the conv module with the projection moved: same parameters, same result as the reference order
    torch.manual_seed(2)
    ref = convs.HyperGsysHGNN(hg, 64, 32).to(cuda_device)
    alt = convs.HyperGsysHGNN(hg, 64, 32, project="edge").to(cuda_device)
    alt.load_state_dict(ref.state_dict())
    X = torch.randn(N, 64, device=cuda_device)
    a, b = ref(X), alt(X)
    assert float((a - b).abs().max() / a.abs().max()) < TOL
    assert ops.projection_order(N, M, 1433, 32) == "vertex" and ops.projection_order(10 ** 6, 4 * 10 ** 5, 256, 256) == "edge"


def test_unigin_and_unigcnii_convs(cuda_device):
    """HyperGsysUinGINConv: (1 + eps) XW + H H^T XW;  HyperGsysUniGCNIIConv: Xi = (1-a) Agg(X) + a X0,
    (1-b) Xi + b W Xi -- forward and input gradient vs fp64."""
    data, hg, V, E, N, M = _setup(cuda_device, "cora", seed=2)
    torch.manual_seed(2)
    Fin, Fout = 48, 32
    degE, degV = hg.degE.double().cpu(), hg.degV.double().cpu()

    gin = convs.HyperGsysUinGINConv(hg, Fin, Fout).to(cuda_device)
    with torch.no_grad():
        gin.eps.fill_(0.25)
    X = torch.randn(N, Fin, device=cuda_device, requires_grad=True)
    out = gin(X)
    out.sum().backward()
    X64 = X.detach().double().cpu().requires_grad_(True)
    Wg = gin.W.weight.detach().double().cpu()
    XW = X64 @ Wg.t()
    want = 1.25 * XW + orc.torch_hgnn_conv(XW, V, E, None, None, None, N, M)
    want.sum().backward()
    assert orc.rel_err(out.detach().cpu().numpy(), want.detach().numpy()) < TOL
    assert orc.rel_err(X.grad.cpu().numpy(), X64.grad.numpy()) < 5e-5

    gcn = convs.HyperGsysUniGCNIIConv(hg, Fout, Fout).to(cuda_device)
    Xa = torch.randn(N, Fout, device=cuda_device, requires_grad=True)
    X0 = torch.randn(N, Fout, device=cuda_device)
    alpha, beta = 0.1, 0.4
    out = gcn(Xa, X0, alpha, beta)
    out.pow(2).sum().backward()
    A64 = Xa.detach().double().cpu().requires_grad_(True)
    Wc = gcn.W.weight.detach().double().cpu()
    Xv = orc.torch_hgnn_conv(A64, V, E, degE, degV, None, N, M)
    Xi = (1 - alpha) * Xv + alpha * X0.double().cpu()
    want = (1 - beta) * Xi + beta * (Xi @ Wc.t())
    want.pow(2).sum().backward()
    assert orc.rel_err(out.detach().cpu().numpy(), want.detach().numpy()) < TOL
    assert orc.rel_err(Xa.grad.cpu().numpy(), A64.grad.numpy()) < 5e-5


def test_training_step_reduces_loss(cuda_device):
    """The epoch protocol of hgsys.py:161-184 (zero_grad -> forward -> nll_loss -> backward -> Adam.step) runs
    and learns on a Cora-shaped graph with planted labels."""
    data, hg, V, E, N, M = _setup(cuda_device, "cora", seed=3)
    torch.manual_seed(3)
    nfeat, ncls = 32, 7
    y = torch.randint(0, ncls, (N,), device=cuda_device)
    X = torch.randn(N, nfeat, device=cuda_device) + Fn.one_hot(y, nfeat).float() * 2.0
    model = convs.HGsysHGNN(None, hg, nfeat, 32, ncls).to(cuda_device)
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4)
    losses = []
    for _ in range(60):
        opt.zero_grad()
        loss = Fn.nll_loss(model(X), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5]) - 0.05   # (dropout 0.6 / 0.6: slow but steady)


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref (the reference compiled in place) is not built")
@pytest.mark.parametrize("shape,F", [("cora", 32), ("pubmed", 64), ("dblp", 128), ("walmart", 32)])
def test_matches_reference_gpu_kernels(shape, F, cuda_device):
    """X3: the un-scaled operator against the reference's own `edge_based_full` lab kernel
    (include/hgnnAgg.cuh:98-131) compiled for sm_100a and run on the same GPU, on C1-C4 shapes."""
    data = synth.make_shape(shape, seed=0)
    hg = HyperGraph(data, cuda_device, data.dataset)
    X = torch.randn(hg.num_nodes, F, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(4))
    ref = orc.ref_lab_gpu(0, hg.ngs, hg.num_edges, hg.group_key, hg.group_start, hg.group_end, hg.H_T_colind, X)
    ours = hgef.UniGNNConv(hg, X)
    scale = ref.abs().max().item()
    # both sides sum in fp32 (the reference with scalar atomics in arbitrary order): 1e-5 of the largest value
    assert ((ours - ref).abs().max().item() / scale) < TOL
    for flags in (_native.HG_FORCE_STREAM, _native.HG_TWO_PASS):
        plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, hg.num_nodes,
                            hg.num_edges)
        got = ops.aggregate(plan, X, flags=flags)
        assert ((got - ref).abs().max().item() / scale) < TOL, flags


def test_partitioned_aggregation_over_nccl_matches_single_gpu(cuda_device):
    """Vertex / hyperedge partitioned path over NCCL (needs >= 2 visible GPUs; skipped on a one-GPU box): every
    rank's block of Y against the single-GPU kernel on the same replicated inputs, through torchrun."""
    import json, os, subprocess, sys
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two visible GPUs")
    world = 4 if ngpu >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for F in ("64", "30"):      # balanced-stage path (F % 4 == 0) and the CSR-kernel path
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.join(root, "tools", "run_partition.py"), "--scale", "0.004", "--F", F, "--iters", "2",
               "--check", "1"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        out = json.loads(line)
        assert out["world"] == world and out["max_rel_err_vs_single_gpu"] < TOL, out


def test_column_parallel_hgnn_matches_single_gpu(cuda_device):
    """ColumnParallelHGNN (first layer split by output columns, all_gather, replicated output layer)
    against HGsysHGNN with the same weights: output, loss, and the gradients of the local column block and of the
    replicated layer.  One rank in-process (the sharding degenerates to the plain model); over NCCL through torchrun
    when the box has two GPUs."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tools", "run_column_parallel.py")
    worlds = [1] + ([2] if torch.cuda.device_count() >= 2 else [])
    for world in worlds:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
               "127.0.0.1", "--master-port", "29519", script] if world > 1 else [sys.executable, script]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert out["world"] == world
        assert out["out_err"] < TOL and out["loss_err"] < 1e-5 and out["grad_w1_block_err"] < 5e-5 and out["grad_w2_err"] < 5e-5, out

