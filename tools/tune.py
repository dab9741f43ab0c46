#!/usr/bin/env python
"""Quick kernel timing for tuning: python tools/tune.py --features 32,128,512 --replicas 64 [--two-pass]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hypergef_b200 as hgef
from hypergef_b200 import ops, synth, _native

ap = argparse.ArgumentParser()
ap.add_argument("--features", default="32,64,128,256,512")
ap.add_argument("--shape", default="pubmed")
ap.add_argument("--replicas", type=int, default=64)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--two-pass", action="store_true")
ap.add_argument("--force-fused", action="store_true")
ap.add_argument("--force-pull", action="store_true")
ap.add_argument("--force-stream", action="store_true")
ap.add_argument("--force-ring", action="store_true")
ap.add_argument("--force-fstream", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--verify", action="store_true", help="compare every result with the stream form's (max abs difference / max abs value)")
ap.add_argument("--sweep", default="", help="semicolon-separated env configs K=V,K=V applied in-process (stream form knobs are read per call)")
args = ap.parse_args()
dev = torch.device("cuda:0")
data = synth.make_shape(args.shape, replicas=args.replicas, seed=0, device=dev)
hg = hgef.HyperGraph(data, dev, data.dataset)
N, M, Z = hg.num_nodes, hg.num_edges, hg.H_T_colind.numel()
plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
W = torch.ones(M, device=dev)
flags = _native.HG_TWO_PASS if args.two_pass else (_native.HG_FORCE_FUSED if args.force_fused else (_native.HG_FORCE_PULL if args.force_pull else (_native.HG_FORCE_STREAM if args.force_stream else (_native.HG_FORCE_RING if args.force_ring else (_native.HG_FORCE_FSTREAM if args.force_fstream else 0)))))
TUNED = set()
KNOBS = ()   # (all knobs go through hg_tune_set: "st_l=64,st_occ=4;..."; lab forms need HGEF_B200_LIB=hypergef_b200/libhgef_b200_lab.so)
for cfg in (args.sweep.split(";") if args.sweep else [""]):
    for k in (KNOBS if args.sweep else ()):
        os.environ.pop(k, None)
    ops.tune(**{k: None for k in TUNED})
    TUNED.clear()
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        ops.tune(**{k: int(v)})
        TUNED.add(k)
    out = []
    for F in [int(f) for f in args.features.split(",")]:
        X = torch.randn(N, F, device=dev); Y = torch.empty(N, F, device=dev)
        for _ in range(3):
            ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, out=Y, flags=flags)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, out=Y, flags=flags)
        b.record(); torch.cuda.synchronize(); plan.check()
        us = a.elapsed_time(b) / args.iters * 1e3
        bad = ""
        if args.verify:
            ops.tune(**{k: None for k in TUNED})
            Y0 = ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, flags=_native.HG_FORCE_STREAM)
            for kv in filter(None, cfg.split(",")):
                k, v = kv.split("="); ops.tune(**{k: int(v)})
            err = float((Y - Y0).abs().max() / Y0.abs().max())
            bad = f" err={err:.1e}" + (" MISMATCH" if not err < 1e-5 else "")
            del Y0
        balg = 8 * F * N + 4 * Z + 12 * M + 4 * N + 4
        if "ring_prof" in TUNED:
            w = plan.debug_words()
            out.append(f"[prof, kclk summed over warps, last launch: control items {w[2]} of which wait-empty {w[3]} wait-deps {w[4]} | worker issue {w[5]} wait-data {w[6]} idle {w[7]}]")
        out.append(f"F={F}: {us:8.1f} us {balg / us / 1e3:7.1f} GB/s ({balg / us / 1e3 / 6536 * 100:4.1f}%){bad}")
        del X, Y
    print(f"[{args.tag} {cfg} {args.shape}x{args.replicas} N={N} Z={Z} heavy={plan.nheavy_edges}] " + " | ".join(out), flush=True)
