#!/usr/bin/env python
"""BASELINE.json configs[2]: UniGNNConv (UniGCN-style, degree-scaled) training forward + backward on a
DBLP-co-authorship-shaped hypergraph, F = 128 -- time of one fwd+bwd pair of the autograd op (two fused
launches: forward, and the true-transpose backward with degV on the gather side)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hypergef_b200 as hgef
from hypergef_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=30)
ap.add_argument("--F", type=int, default=128)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda:0")
data = synth.make_shape("dblp", replicas=args.replicas, seed=0, device=dev)
hg = hgef.HyperGraph(data, dev, data.dataset)
N, M, Z, F = hg.num_nodes, hg.num_edges, int(hg.H_T_colind.numel()), args.F
X = torch.randn(N, F, device=dev, requires_grad=True)
G = torch.randn(N, F, device=dev)
def step():
    X.grad = None
    hgef.UniGNNConvdeg(hg, X, hg.degE, hg.degV).backward(G)
for _ in range(3):
    step()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.iters):
    step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.iters
balg = 2 * (8 * F * N + 4 * Z + 8 * M + 4 * N + 4)
print(json.dumps({"shape": f"dblp x{args.replicas}", "N": N, "M": M, "nnz": Z, "F": F, "fwd_bwd_ms": ms,
                  "algorithmic_GBps": balg / ms / 1e6, "frac_of_measured_peak": balg / ms / 1e6 / 6536}), flush=True)
