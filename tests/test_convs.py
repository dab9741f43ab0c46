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
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tools", "run_partition.py"), "--scale", "0.004", "--F", "64", "--iters", "2",
           "--check", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == world and out["max_rel_err_vs_single_gpu"] < TOL, out
