#!/usr/bin/env python
"""Timing experiment: the two gather phases with the plain warp-per-row kernel (hg_edge_reduce)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hypergef_b200 as hgef
from hypergef_b200 import synth
from hypergef_b200.partition import CudaBackend
dev = torch.device("cuda:0")
data = synth.make_shape("pubmed", replicas=64, seed=0, device=dev)
hg = hgef.HyperGraph(data, dev, "pubmed")
be = CudaBackend(dev, 40)
N, M, Z = hg.num_nodes, hg.num_edges, hg.H_T_colind.numel()
for F in (32, 64, 128, 256, 512):
    X = torch.randn(N, F, device=dev)
    def run():
        P = be.edge_reduce(hg.H_T_csrptr, hg.H_T_colind, X, None)
        return be.edge_reduce(hg.H_csrptr, hg.H_colind, P, None)
    for _ in range(3): run()
    ts = []
    for fn in (lambda: be.edge_reduce(hg.H_T_csrptr, hg.H_T_colind, X, None), None, run):
        if fn is None:
            P = be.edge_reduce(hg.H_T_csrptr, hg.H_T_colind, X, None)
            fn = lambda: be.edge_reduce(hg.H_csrptr, hg.H_colind, P, None)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 10 * 1e3)
    balg = 8 * F * N + 4 * Z + 12 * M + 4 * N + 4
    print(f"F={F}: A {ts[0]:.1f} us  B {ts[1]:.1f} us  both {ts[2]:.1f} us  ({balg / ts[2] / 1e3:.0f} GB/s alg, {balg / ts[2] / 1e3 / 6536 * 100:.1f}%)", flush=True)
