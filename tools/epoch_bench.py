#!/usr/bin/env python
"""HGNN epoch ms (BASELINE.json configs[0]): 2-layer HGNN, fp32, Cora-shaped synthetic hypergraph
(2708 vertices, 1579 hyperedges, 1433 -> 32 -> 7), protocol of HyperGsys/hgsys.py:161-184
(10 warm-up steps, then `epochs` steps of zero_grad -> forward -> nll_loss -> backward -> Adam.step,
synchronise before/after, mean per step).  GPU arm = hypergef_b200 convs (optionally CUDA-graphed);
CPU arm = the PyG-equivalent pure-torch conv (model/pygnn/hgnn.py:25-38) on the host cores."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as Fn
import hypergef_b200 as hgef
from hypergef_b200 import convs, synth

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="cora")
ap.add_argument("--nfeat", type=int, default=1433)
ap.add_argument("--nhid", type=int, default=32)
ap.add_argument("--nclass", type=int, default=7)
ap.add_argument("--epochs", type=int, default=200)
ap.add_argument("--cpu-epochs", type=int, default=20)
args = ap.parse_args()
torch.manual_seed(1)                                     # hgsys.py:56,76-77
data = synth.make_shape(args.shape, seed=0, num_feat=args.nfeat)
N = data.num_nodes
y = data.y % args.nclass
out = {"config": f"{args.shape}-shaped, 2-layer HGNN {args.nfeat}->{args.nhid}->{args.nclass}, dropout 0.6/0.6, Adam lr 0.01 wd 5e-4"}


def run(model, X, y, epochs, sync):
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4)      # hgsys.py:136
    def step():
        model.train(); opt.zero_grad()
        loss = Fn.nll_loss(model(X), y); loss.backward(); opt.step()
        return loss
    for _ in range(10):
        step()
    sync(); t0 = time.perf_counter()
    for _ in range(epochs):
        loss = step()
    sync()
    return (time.perf_counter() - t0) / epochs * 1e3, float(loss)


if torch.cuda.is_available():
    dev = torch.device("cuda:0")
    hg = hgef.HyperGraph(data, dev, data.dataset)
    model = convs.HGsysHGNN(None, hg, args.nfeat, args.nhid, args.nclass).to(dev)
    ms, loss = run(model, data.x.to(dev), y.to(dev), args.epochs, torch.cuda.synchronize)
    out["gpu_ms_per_epoch"], out["gpu_final_loss"] = ms, loss
    # the same step captured once in a CUDA graph and replayed (the launch-bound regime of small graphs)
    try:
        torch.manual_seed(1)
        gmodel = convs.HGsysHGNN(None, hg, args.nfeat, args.nhid, args.nclass).to(dev)
        gopt = torch.optim.Adam(gmodel.parameters(), lr=0.01, weight_decay=5e-4, capturable=True)
        Xd, yd = data.x.to(dev), y.to(dev)
        gmodel.train()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                gopt.zero_grad(set_to_none=True)
                Fn.nll_loss(gmodel(Xd), yd).backward()
                gopt.step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        gopt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            gloss = Fn.nll_loss(gmodel(Xd), yd)
            gloss.backward()
            gopt.step()
        for _ in range(10):
            graph.replay()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(args.epochs):
            graph.replay()
        torch.cuda.synchronize()
        out["gpu_graph_ms_per_epoch"] = (time.perf_counter() - t0) / args.epochs * 1e3
        out["gpu_graph_final_loss"] = float(gloss)
    except Exception as exc:  # report, do not hide
        out["gpu_graph_error"] = repr(exc)[:300]


class CpuConv(nn.Module):                                # model/pygnn/hgnn.py:10-38 restated
    def __init__(self, V, E, degE, degV, n, m, cin, cout):
        super().__init__()
        self.W = nn.Linear(cin, cout, bias=False)
        self.V, self.E, self.degE, self.degV, self.n, self.m = V, E, degE, degV, n, m
    def forward(self, X):
        from oracle import oracle as orc
        return orc.torch_hgnn_conv(self.W(X), self.V, self.E, self.degE, self.degV, None, self.n, self.m)


class CpuHGNN(nn.Module):                                # model/gnn.py:31-70
    def __init__(self, mk, nfeat, nhid, nclass):
        super().__init__()
        self.c1, self.c2 = mk(nfeat, nhid), mk(nhid, nclass)
        self.d0, self.d1 = nn.Dropout(0.6), nn.Dropout(0.6)
    def forward(self, X):
        X = self.d1(torch.relu(self.c1(self.d0(X))))
        return Fn.log_softmax(self.c2(X), dim=1)


from oracle import oracle as orc
V, E, M, Z = orc.split_edge_index(data.edge_index, N)
H, _ = orc.scipy_incidence(V.numpy(), E.numpy(), N, M)
degV, degE = orc.scipy_degrees(H)
cpu = CpuHGNN(lambda a, b: CpuConv(V, E, degE, degV, N, M, a, b), args.nfeat, args.nhid, args.nclass)
ms, loss = run(cpu, data.x, y, args.cpu_epochs, lambda: None)
out["cpu_ms_per_epoch"], out["cpu_threads"], out["cpu_final_loss"] = ms, torch.get_num_threads(), loss
print(json.dumps(out), flush=True)
