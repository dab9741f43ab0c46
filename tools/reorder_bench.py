#!/usr/bin/env python
"""What locality reordering is worth on this path: the bench graph (Pubmed-shaped x replicas) in its natural numbering,
with vertex and hyperedge ids shuffled (no locality left), and after reverse Cuthill-McKee on the shuffled graph
(hypergef_b200.reorder).  Prints the aggregation time per feature length for each."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hypergef_b200 as hgef
from hypergef_b200 import ops, reorder, synth

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=64)
ap.add_argument("--features", default="64,128,256,512")
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda:0")
data = synth.make_shape("pubmed", replicas=args.replicas, seed=0)
N, M = data.num_nodes, data.num_hyperedges
g = torch.Generator().manual_seed(7)
shuf = reorder.permute_data(data, torch.randperm(N, generator=g), torch.randperm(M, generator=g))
t0 = time.time()
rcm, _, _ = reorder.reorder(shuf)
t_rcm = time.time() - t0
out = {"N": N, "M": M, "rcm_host_seconds": round(t_rcm, 2), "rows": []}
for name, d in (("natural", data), ("shuffled", shuf), ("rcm_of_shuffled", rcm)):
    hg = hgef.HyperGraph(d, dev, "pubmed")
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    W = torch.ones(M, device=dev)
    row = {"order": name, "mean_hyperedge_span": round(reorder.mean_span(d.edge_index, N))}
    for F in [int(f) for f in args.features.split(",")]:
        X = torch.randn(N, F, device=dev)
        Y = torch.empty_like(X)
        for _ in range(3):
            ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, out=Y)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, out=Y)
        b.record()
        torch.cuda.synchronize()
        row[f"F{F}_us"] = round(a.elapsed_time(b) / args.iters * 1e3, 1)
        del X, Y
    out["rows"].append(row)
    del hg, plan
    ops.clear_plan_cache()
    torch.cuda.empty_cache()
print(json.dumps(out))
