#!/usr/bin/env python
"""Small stream-form workload (every launch shape, checked against the two-pass kernels) -- sized for a run under
`compute-sanitizer --tool memcheck` where that tool is available (it is closed on the round's GPU pool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hypergef_b200 as hgef
from hypergef_b200 import ops, synth, _native

dev = torch.device("cuda:0")
for shape, reps in (("pubmed", 1), ("walmart", 1)):
    data = synth.make_shape(shape, replicas=reps, seed=1)
    hg = hgef.HyperGraph(data, dev, data.dataset)
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, hg.num_nodes, hg.num_edges)
    W = torch.rand(hg.num_edges, device=dev) + 0.5
    for F in (4, 32, 100, 128, 256, 640):
        X = torch.randn(hg.num_nodes, F, device=dev)
        ref = ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, flags=_native.HG_TWO_PASS)
        for knobs in ({}, {"st_pdl": 0}, {"st_l": 16}, {"st_pipe": 0}):
            ops.tune(st_pdl=None, st_l=None, st_pipe=None)
            ops.tune(**knobs)
            out = ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, flags=_native.HG_FORCE_STREAM)
            out_t = ops.aggregate(plan, X, s1=hg.degE, s2=W, a_in=hg.degV, flags=_native.HG_FORCE_STREAM)
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            assert err < 1e-5, (shape, F, knobs, err)
        plan.check()
    print("ok", shape, hg.num_nodes, plan.nheavy_edges, flush=True)
