#!/usr/bin/env python
"""Partitioned multi-GPU aggregation on real GPUs (NCCL):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_partition.py
Checks the distributed result against the single-GPU kernel on rank 0 and reports time + exchange bytes."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import hypergef_b200 as hgef
from hypergef_b200 import ops, synth
from hypergef_b200.partition import CudaBackend, PartitionedAggregator, build_partition

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="c5")
ap.add_argument("--scale", type=float, default=0.02, help="fraction of the named shape's N and E (c5 = 50M x 10M)")
ap.add_argument("--replicas", type=int, default=1)
ap.add_argument("--F", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--single-stage-a", action="store_true", help="one stage A over every local hyperedge (no overlap of the exchange)")
ap.add_argument("--mode", default="partition", choices=["partition", "colshard"],
                help="partition: vertex/hyperedge blocks + NCCL boundary exchange; colshard: every rank owns F/world feature columns of the whole graph (no collective)")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
shape = synth.SHAPES[args.shape]
if args.scale != 1.0:
    import dataclasses
    shape = dataclasses.replace(shape, num_nodes=int(shape.num_nodes * args.scale), num_edges=int(shape.num_edges * args.scale))
data = synth.make_shape(args.shape, replicas=args.replicas, seed=0, device=dev, shape=shape)   # same seed on every rank
hg = hgef.HyperGraph(data, dev, "synthetic", ngs=shape.ngs)
N, M, F = hg.num_nodes, hg.num_edges, args.F
if args.mode == "colshard":
    Fl = F // world
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    Xl = torch.empty(N, Fl, device=dev)
    for i in range(0, N, 1 << 22):
        Xl[i:i + (1 << 22)].normal_(generator=gen)
    Yl = torch.empty_like(Xl)
    W = torch.ones(M, device=dev)
    for _ in range(3):
        ops.aggregate(plan, Xl, s1=hg.degE, s2=W, a_out=hg.degV, out=Yl)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        ops.aggregate(plan, Xl, s1=hg.degE, s2=W, a_out=hg.degV, out=Yl)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    Z = int(hg.H_T_colind.numel())
    balg = 8 * F * N + world * (4 * Z + 12 * M + 4 * N + 4)
    if rank == 0:
        print(json.dumps({"mode": "colshard", "world": world, "N": N, "M": M, "nnz": Z, "F": F, "F_per_rank": Fl,
                          "ms": ms.item(), "algorithmic_GBps": balg / ms.item() / 1e6}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)
info = build_partition(hg.H_T_csrptr, hg.H_T_colind, N, M, world, rank)
agg = PartitionedAggregator(info, CudaBackend(dev, shape.ngs), split_stage_a=not args.single_stage_a)
gen = torch.Generator(device=dev).manual_seed(5)
Xfull = torch.randn(N, F, device=dev, generator=gen) if args.check else None
Xl = (Xfull[info.v_start:info.v_end].contiguous() if args.check else
      torch.randn(info.num_local, F, device=dev, generator=gen))
degE, degV = hg.degE.reshape(-1), hg.degV.reshape(-1)
dl = degV[info.v_start:info.v_end].contiguous()
W = torch.ones(M, device=dev)
Y = agg.forward(Xl, s1=degE, s2=W, a_out=dl)
torch.cuda.synchronize()
err = None
if args.check:
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    want = ops.aggregate(plan, Xfull, s1=degE, s2=W, a_out=degV)[info.v_start:info.v_end]
    err = ((Y - want).abs().max() / want.abs().max()).item()
for _ in range(3):
    agg.forward(Xl, s1=degE, s2=W, a_out=dl)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
agg.bytes_exchanged = 0
a.record()
for _ in range(args.iters):
    agg.forward(Xl, s1=degE, s2=W, a_out=dl)
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / args.iters], device=dev)
errt = torch.tensor([err if err is not None else 0.0], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(errt, op=dist.ReduceOp.MAX)
Z = int(hg.H_T_colind.numel())
balg = 8 * F * N + 4 * Z + 12 * M + 4 * N + 4
if rank == 0:
    print(json.dumps({"world": world, "N": N, "M": M, "nnz": Z, "F": F, "ms": ms.item(), "algorithmic_GBps": balg / ms.item() / 1e6,
                      "boundary_hyperedges": info.num_boundary_total, "interior_rank0": int(info.int_edges.numel()),
                      "exchange_bytes_per_call_rank0": agg.bytes_exchanged // args.iters,
                      "max_rel_err_vs_single_gpu": errt.item() if args.check else None}), flush=True)
if world > 1:
    dist.destroy_process_group()
