#!/usr/bin/env python
"""bench.py -- fused hypergraph aggregation throughput (BASELINE.json configs[1]) and HGNN epoch ms.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, C-ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (host cores)

N = 1 (``config.workload``): the standalone fused-aggregation sweep, feature length 32..512, on a synthetic
Pubmed-shaped hypergraph (19717 vertices / 7963 hyperedges per replica, ngs=40) stacked block-diagonally
``--replicas`` times so that X and Y are each larger than the 126 MB L2 at every F (literal Pubmed is
L2-resident on a B200).  One STEP = one pass of the hot path over the batch: the HGNN
aggregation ``Y = degV.H.(degE*W).H^T.X`` once per feature length of the sweep.  ``value`` = algorithmic
GB/s = sum_F B_alg(F) / step time with
    B_alg(F) = 8*F*N + 4*Z + 12*E + 4*N + 4   bytes  (every array once)
and inputs resident in HBM.  ``e2e`` = the same metric through the host-buffer API (pinned host X -> device
-> aggregation -> host Y inside the timed region).  Extra blocks on the same line: ``roofline``,
``cpu_baseline``, ``same_gpu_baselines`` (cuSPARSE 2xSpMM, the reference's own kernels compiled for sm_100a,
this repo's two-pass form), ``epoch`` (HGNN epoch ms, BASELINE configs[0]) and ``c5`` (the C5-shaped
graph of the multi-GPU runs on this one GPU, for the strong-scaling baseline).

N > 1: STRONG scaling of ONE aggregation on the C5-shaped graph (BASELINE configs[4]; ``--c5-scale`` of
50 M x 10 M, F = 256), two ways, both in the JSON: (a) feature-column sharding (every rank owns F/N columns of
the replicated graph; no data-path collective), (b) vertex / hyperedge partition with the NCCL boundary
exchange inside the timed region.  ``value`` = algorithmic GB/s of the WHOLE problem / the faster path's time
(max over ranks); ``config.parallelism`` names it.  ``e2e`` = path (a) from pinned host buffers.  ``epoch`` =
the column-parallel 2-layer HGNN (all_gather of the hidden activations).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FEATURES = (32, 64, 128, 256, 512)
METRIC = "fused_aggregation_algorithmic_GBps"
UNIT = "GB/s"


def b_alg(N, E, Z, F):
    return 8 * F * N + 4 * Z + 12 * E + 4 * N + 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled from a thread
    every few ms (the timed region of the default run is ~50 ms, shorter than one nvidia-smi period); if NVML
    is unavailable the nvidia-smi loop of the recipe is used instead."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.stop = None, None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample(self):
        """One NVML reading, taken by the caller while the device is still executing the timed steps."""
        if self.nvml is None:
            return
        n = self.nvml
        try:
            r = (getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons)(self.handle)
            bits = (getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4))
            if not hasattr(self, "_max"):          # (NVML calls take milliseconds: the maximum is read once)
                self._max = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            self.rows.append([str(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)), str(self._max), ""] +
                             ["Active" if r & b else "Not Active" for b in bits])
        except Exception:
            pass

    def _poll_nvml(self):
        n = self.nvml
        bits = ((getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap"))
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while True:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                r = get_reasons(self.handle)
                self.rows.append([str(sm), str(mx), ""] + ["Active" if r & b else "Not Active" for b, _ in bits])
            except Exception:
                break
            if self.stop.wait(0.004):
                break

    def __enter__(self):
        if self.nvml is not None:
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.nvml is not None:
            self.stop.set()
            self.t.join(timeout=2)
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(self.NAMES, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------- workloads
def c5_shape(scale):
    import dataclasses
    from hypergef_b200 import synth
    shp = synth.SHAPES["c5"]
    if scale != 1.0:
        shp = dataclasses.replace(shp, num_nodes=int(shp.num_nodes * scale), num_edges=int(shp.num_edges * scale))
    return shp


def host_threads():
    """All host cores for the CPU arms: torchrun exports OMP_NUM_THREADS=1, which would make the reference run
    on one core (and every ratio against it 8-16x too good)."""
    n = os.cpu_count() or 1
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)
    return torch.get_num_threads()


# ----------------------------------------------------------------------------- reference arm
def cpu_conv_workload(shape_name, replicas, features, seed=0, shape=None):
    """Graph + features for the reference's PyG-equivalent CPU conv (model/pygnn/hgnn.py:30-37)."""
    from hypergef_b200 import synth
    from oracle import oracle as orc
    data = synth.make_shape(shape_name, replicas=replicas, seed=seed, shape=shape)
    N = data.num_nodes
    V, E, M, Z = orc.split_edge_index(data.edge_index, N)
    H, _ = orc.scipy_incidence(V.numpy(), E.numpy(), N, M)
    degV, degE = orc.scipy_degrees(H)
    W = torch.ones(M)
    Xs = {F: torch.randn(N, F, generator=torch.Generator().manual_seed(F)) for F in features}
    bytes_step = sum(b_alg(N, M, Z, F) for F in features)

    def step():
        out = None
        for F in features:
            out = orc.torch_hgnn_conv(Xs[F], V, E, degE, degV, W, N, M)
        return out
    return step, bytes_step, dict(N=N, E=M, Z=int(Z))


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference(args, rank):
    """The reference's CPU path on the box's host cores: same workload, steps and warm-up as our arm at N = 1; at
    N > 1 (our arm: one aggregation on the C5-shaped graph) a bounded sample of that graph."""
    if rank != 0:
        return
    cores = host_threads()
    warmup = max(args.warmup, 3)
    if args.gpus == 1:
        features = tuple(args.features)
        step, bytes_step, dims = cpu_conv_workload("pubmed", args.replicas, features)
        cfg = workload_config(args)
        sample = (f"the whole N=1 workload: pubmed-shaped x{args.replicas} replicas (N={dims['N']}, E={dims['E']}, "
                  f"nnz={dims['Z']}), F sweep {list(features)}")
    else:
        scale = min(args.c5_scale, args.ref_c5_scale)
        step, bytes_step, dims = cpu_conv_workload("c5", 1, (args.c5_feature,), shape=c5_shape(scale))
        cfg = c5_config(args)
        sample = (f"bounded sample of the N>1 workload: C5-shaped graph at scale {scale} (N={dims['N']}, E={dims['E']}, "
                  f"nnz={dims['Z']}) instead of {args.c5_scale}, F={args.c5_feature}")
    sample += ("; pure-torch restatement of model/pygnn/hgnn.py:30-37 (index_select + index_add_; torch_scatter / PyG "
               "are not installed), a CPU PORT of the reference's PyG back-end, not its CUDA kernels")
    sec = time_cpu(step, args.steps, warmup)
    val = bytes_step / sec / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"pubmed-shaped hypergraph x{args.replicas} block-diagonal replicas, fused HGNN aggregation, "
                        f"F sweep {list(args.features)}",
            "shape": "pubmed (19717 vertices, 7963 hyperedges per replica)", "replicas": args.replicas,
            "ngs": 40, "features": list(args.features),
            "l2": "X and Y of every launch are each > 126 MB L2 (inputs larger than L2; no flush)",
            "parallelism": "single GPU"}


def c5_config(args):
    shp = c5_shape(args.c5_scale)
    return {"workload": f"C5-shaped hypergraph (BASELINE configs[4]) at scale {args.c5_scale}: {shp.num_nodes} vertices, "
                        f"{shp.num_edges} hyperedges, mean size 10, members inside a 65536-vertex window with "
                        f"probability 0.9; ONE fused HGNN aggregation at F={args.c5_feature}",
            "shape": "c5", "scale": args.c5_scale, "ngs": shp.ngs, "features": [args.c5_feature],
            "l2": "X and Y are far larger than the 126 MB L2 (no flush)"}


# ----------------------------------------------------------------------------- helpers of our arm
def timed_loop(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    barrier()
    return a.elapsed_time(b) / steps


def max_over_ranks(ms, dev, world):
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def pcie_probe(dev, world, mb=512):
    """Pinned host <-> device copy rate of every rank AT THE SAME TIME (both directions together): the ceiling
    of the e2e leg on this box, where every GPU hangs off the same host memory."""
    import torch.distributed as dist
    n = mb * 1024 * 1024 // 4
    h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
    d_in, d_out = torch.empty(n, device=dev), torch.empty(n, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        s1.synchronize(); s2.synchronize()
    both()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        both()
    sec = (time.perf_counter() - t0) / 3
    gbs = torch.tensor([2 * n * 4 / sec / 1e9], device=dev, dtype=torch.float64)
    tot = gbs.clone()
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    return {"per_rank_GBps_both_directions": gbs.item(), "aggregate_GBps": tot.item(), "buffer_MB": mb}


def epoch_block(dev, rank, world, args):
    """HGNN epoch ms (BASELINE configs[0]): 2-layer HGNN fp32 on the Cora-shaped hypergraph (2708 vertices, 1579
    hyperedges, 1433 -> 32 -> 7), protocol of hgsys.py:161-184 (10 warm-up steps, then `epochs` steps of zero_grad ->
    forward -> nll_loss -> backward -> Adam.step, synchronise before/after, mean per step; dropout 0.6 / 0.6, Adam
    lr 0.01 wd 5e-4, seed 1).  N = 1: eager, and the same step captured in a CUDA graph, with the PyG-equivalent CPU
    conv beside it.  N > 1: the column-parallel model (all_gather of the hidden activations), eager."""
    import torch.nn.functional as Fn
    import hypergef_b200 as hgef
    from hypergef_b200 import convs, synth
    import torch.distributed as dist
    nfeat, nhid, ncls, epochs = 1433, 32, 7, args.epochs
    torch.manual_seed(1)
    data = synth.make_shape("cora", seed=0, num_feat=nfeat)
    hg = hgef.HyperGraph(data, dev, data.dataset)
    X, y = data.x.to(dev), (data.y % ncls).to(dev)
    out = {"config": f"cora-shaped (2708 x 1579), 2-layer HGNN {nfeat}->{nhid}->{ncls}, dropout 0.6/0.6, Adam lr 0.01 wd 5e-4, "
                     f"{epochs} epochs after 10 warm-up (hgsys.py:161-184)", "n_gpus": world}

    def run(model, opt, epochs):
        def step():
            model.train(); opt.zero_grad()
            loss = Fn.nll_loss(model(X), y); loss.backward(); opt.step()
            return loss
        for _ in range(10):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(epochs):
            loss = step()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / epochs * 1e3, float(loss)

    if world == 1:
        model = convs.HGsysHGNN(None, hg, nfeat, nhid, ncls).to(dev)
        ms, loss = run(model, torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4), epochs)
        out["gpu_ms_per_epoch"], out["gpu_final_loss"] = ms, loss
        try:    # the same step captured once in a CUDA graph and replayed (small graphs are launch-bound)
            torch.manual_seed(1)
            gm = convs.HGsysHGNN(None, hg, nfeat, nhid, ncls).to(dev)
            gopt = torch.optim.Adam(gm.parameters(), lr=0.01, weight_decay=5e-4, capturable=True)
            gm.train()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    gopt.zero_grad(set_to_none=True)
                    Fn.nll_loss(gm(X), y).backward()
                    gopt.step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            gopt.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                gloss = Fn.nll_loss(gm(X), y)
                gloss.backward()
                gopt.step()
            for _ in range(10):
                graph.replay()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(epochs):
                graph.replay()
            torch.cuda.synchronize()
            out["gpu_graph_ms_per_epoch"] = (time.perf_counter() - t0) / epochs * 1e3
        except Exception as exc:  # report, do not hide
            out["gpu_graph_error"] = repr(exc)[:300]
        if not args.no_cpu_baseline:
            out.update(cpu_epoch(data, nfeat, nhid, ncls, args.cpu_epochs))
    else:
        model = convs.ColumnParallelHGNN(hg, nfeat, nhid, ncls).to(dev)
        # identical replicated output layer on every rank (same seed); first-layer blocks differ by construction
        ms, loss = run(model, torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4), epochs)
        out["gpu_ms_per_epoch"] = max_over_ranks(ms, dev, world)
        out["gpu_final_loss"] = loss
        out["parallelism"] = f"feature-column parallel first layer ({nhid}/{world} hidden columns per rank), all_gather, replicated output layer"
    return out


def cpu_epoch(data, nfeat, nhid, ncls, epochs):
    """The reference's pyg back-end restated in pure torch (model/pygnn/hgnn.py:10-38, model/gnn.py:31-70) on the host."""
    import torch.nn as nn
    import torch.nn.functional as Fn
    from oracle import oracle as orc
    N = data.num_nodes
    V, E, M, Z = orc.split_edge_index(data.edge_index, N)
    H, _ = orc.scipy_incidence(V.numpy(), E.numpy(), N, M)
    degV, degE = orc.scipy_degrees(H)

    class Conv(nn.Module):
        def __init__(self, cin, cout):
            super().__init__()
            self.W = nn.Linear(cin, cout, bias=False)

        def forward(self, X):
            return orc.torch_hgnn_conv(self.W(X), V, E, degE, degV, None, N, M)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.c1, self.c2 = Conv(nfeat, nhid), Conv(nhid, ncls)
            self.d0, self.d1 = nn.Dropout(0.6), nn.Dropout(0.6)

        def forward(self, X):
            return Fn.log_softmax(self.c2(self.d1(torch.relu(self.c1(self.d0(X))))), dim=1)
    cores = host_threads()
    torch.manual_seed(1)
    net = Net()
    opt = torch.optim.Adam(net.parameters(), lr=0.01, weight_decay=5e-4)
    X, y = data.x, data.y % ncls

    def step():
        net.train(); opt.zero_grad()
        loss = Fn.nll_loss(net(X), y); loss.backward(); opt.step()
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(epochs):
        step()
    return {"cpu_ms_per_epoch": (time.perf_counter() - t0) / epochs * 1e3, "cpu_threads": cores, "cpu_epochs": epochs}


def c5_block(dev, rank, world, args, barrier):
    """ONE aggregation on the C5-shaped graph: column-sharded over the ranks (and, N > 1, vertex / hyperedge
    partitioned with the NCCL boundary exchange in the timed region).  Returns the block and the objects the
    host-buffer leg re-uses."""
    import hypergef_b200 as hgef
    from hypergef_b200 import ops, synth
    from hypergef_b200.partition import CudaBackend, PartitionedAggregator, build_partition
    shp = c5_shape(args.c5_scale)
    F = args.c5_feature
    data = synth.make_shape("c5", seed=0, device=dev, shape=shp)      # same seed on every rank: replicated structure
    hg = hgef.HyperGraph(data, dev, "synthetic", ngs=shp.ngs)
    del data
    N, M, Z = hg.num_nodes, hg.num_edges, int(hg.H_T_colind.numel())
    W = torch.ones(M, device=dev)
    full_bytes = b_alg(N, M, Z, F)
    steps, warm = max(3, min(args.steps, 10)), 3
    blk = {"N": N, "E": M, "nnz": Z, "F": F, "algorithmic_bytes": full_bytes}

    # (a) feature columns: every rank F / world columns of the whole graph, no collective
    Fl = F // world
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    Xl = torch.empty(N, Fl, device=dev)
    for i in range(0, N, 1 << 22):
        Xl[i:i + (1 << 22)].normal_(generator=gen)
    Yl = torch.empty_like(Xl)
    ms = timed_loop(lambda: ops.aggregate(plan, Xl, s1=hg.degE, s2=W, a_out=hg.degV, out=Yl), steps, warm, barrier)
    ms = max_over_ranks(ms, dev, world)
    blk["column_sharded"] = {"ms": ms, "algorithmic_GBps": full_bytes / ms / 1e6, "columns_per_rank": Fl,
                             "collective": "none (columns are independent, hgnnaggr_cuda.cu:21,34,44); the graph and "
                                           "balancer arrays are replicated"}
    keep = dict(hg=hg, plan=plan, W=W, N=N, Fl=Fl, full_bytes=full_bytes)
    if world > 1:
        # (b) vertex / hyperedge partition: rank r owns a contiguous vertex block; boundary hyperedge features cross NVLink
        del Xl, Yl
        torch.cuda.empty_cache()
        info = build_partition(hg.H_T_csrptr, hg.H_T_colind, N, M, world, rank)
        agg = PartitionedAggregator(info, CudaBackend(dev, shp.ngs))
        gen = torch.Generator(device=dev).manual_seed(5)
        Xp = torch.empty(info.num_local, F, device=dev)
        for i in range(0, info.num_local, 1 << 21):
            Xp[i:i + (1 << 21)].normal_(generator=gen)
        degE, degV = hg.degE.reshape(-1), hg.degV.reshape(-1)
        dl = degV[info.v_start:info.v_end].contiguous()
        agg.bytes_exchanged = 0
        msp = timed_loop(lambda: agg.forward(Xp, s1=degE, s2=W, a_out=dl), steps, warm, barrier)
        msp = max_over_ranks(msp, dev, world)
        per_call = agg.bytes_exchanged / (steps + warm)
        blk["partitioned"] = {"ms": msp, "algorithmic_GBps": full_bytes / msp / 1e6,
                              "boundary_hyperedges": info.num_boundary_total, "boundary_fraction": info.num_boundary_total / M,
                              "exchanged_bytes_per_rank_per_call": per_call,
                              "collective": "2 x NCCL all_to_all_single of boundary hyperedge rows (partials to owners, completed rows "
                                            "back), on a side stream while stage A of the interior hyperedges computes; both stages are the "
                                            "balanced stream kernels over local plans (hg_plan_edge_reduce / hg_plan_edge_scatter)"}
        del Xp, agg
        torch.cuda.empty_cache()
    return blk, keep


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    if world == 1:
        return run_single(args, dev, barrier)
    return run_multi(args, dev, rank, world, barrier)


def run_single(args, dev, barrier):
    import hypergef_b200 as hgef
    from hypergef_b200 import ops, synth
    features = tuple(args.features)
    data = synth.make_shape("pubmed", replicas=args.replicas, seed=0, device=dev)
    hg = hgef.HyperGraph(data, dev, "pubmed")
    N, M, Z = hg.num_nodes, hg.num_edges, int(hg.H_T_colind.numel())
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    plan.reserve(max(features))
    W = torch.ones(M, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1000)
    Xs = {F: torch.randn(N, F, device=dev, generator=gen) for F in features}
    Ys = {F: torch.empty(N, F, device=dev) for F in features}
    bytes_f = {F: b_alg(N, M, Z, F) for F in features}
    bytes_step = sum(bytes_f.values())
    warmup = max(args.warmup, 3)

    def call(F):
        ops.aggregate(plan, Xs[F], s1=hg.degE, s2=W, a_out=hg.degV, out=Ys[F])
    for _ in range(warmup):
        for F in features:
            call(F)
    launches0 = plan.kernels_launched()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in features]
          for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(dev.index) as clocks:
        t0.record()
        for k in range(args.steps):
            for j, F in enumerate(features):
                ev[k][j][0].record()
                call(F)
                ev[k][j][1].record()
        t1.record()
        while not t1.query():          # the host is far ahead of the device: read the clocks while it executes
            clocks.sample()
            time.sleep(0.002)
        barrier()
    kernels = plan.kernels_launched() - launches0
    ms_step = t0.elapsed_time(t1) / args.steps
    per_f_ms = {F: float(np.mean([ev[k][j][0].elapsed_time(ev[k][j][1]) for k in range(args.steps)]))
                for j, F in enumerate(features)}
    value = bytes_step / (ms_step * 1e-3) / 1e9

    # ---- end to end: host buffers, copies inside the timed region -------------------------
    Fsum = sum(features)
    hx = torch.empty(N * Fsum, dtype=torch.float32).pin_memory()
    hy = torch.empty(N * Fsum, dtype=torch.float32).pin_memory()
    hx.normal_(generator=torch.Generator().manual_seed(7))
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    # the host-buffer API a numpy / CPU-torch caller uses; the sweep's five calls are submitted
    # back to back so uploads, launches and downloads overlap, then the step waits for all results
    pipe = ops.HostPipeline(plan)
    offs, o = {}, 0
    for F in features:
        offs[F] = o
        o += N * F

    def e2e_step():
        for F in features:
            a, b = offs[F], offs[F] + N * F
            pipe.submit(hx[a:b].view(N, F), hy[a:b].view(N, F), s1=hg.degE, s2=W, a_out=hg.degV)
        pipe.wait()
    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    e2e_val = bytes_step / (e2e_ms * 1e-3) / 1e9
    io_bytes = sum(4 * F * N for F in features)
    del hx, hy, pipe

    peak, peak_src = peaks()
    kern_ms = sum(per_f_ms.values())
    achieved = bytes_step / (kern_ms * 1e-3) / 1e9
    # gather_scatter_GBps: the reference's own work unit beside the algorithmic figure (B_gs = 8 F Z + 8 Z,
    # every member row gathered once and scattered once; include/spmm/spmm.cuh:663,717 counts 4 F nnz per two-step)
    sweep = [{"F": F, "us": per_f_ms[F] * 1e3, "algorithmic_GBps": bytes_f[F] / (per_f_ms[F] * 1e-3) / 1e9,
              "frac_of_measured_peak": bytes_f[F] / (per_f_ms[F] * 1e-3) / 1e9 / peak,
              "gather_scatter_GBps": (8 * F * Z + 8 * Z) / (per_f_ms[F] * 1e-3) / 1e9} for F in features]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_step")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args), N=N, E=M, nnz=Z,
                           heavy_hyperedges=plan.nheavy_edges, segments=plan.nseg),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "hg_aggr_forward per F = stream form: stream_kernel stage A (X -> Xe over balancer segments) + "
                                   "stage B (Xe -> Y over vertices; a programmatic dependent launch of A), two launches per call, "
                                   "CUDA events around each C-ABI call",
                         "algorithmic_bytes_per_step": bytes_step, "sweep": sweep},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes,
                    "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": kernels,
            "clocks": clocks.summary()}
    if not args.no_extras:
        line["same_gpu_baselines"] = same_gpu_baselines(hg, plan, Xs, Ys, W, features, bytes_f, peak)
        line["literal_shapes"] = literal_shapes(dev)
        try:
            line["layer"] = layer_block(hg, W)
        except Exception as exc:    # an extra must not cost the line
            line["layer"] = {"error": repr(exc)[:300]}
    if not args.no_cpu_baseline:
        cores = host_threads()
        step, b_cpu, dims = cpu_conv_workload("pubmed", args.replicas, features)
        sec = time_cpu(step, 2, 1)
        line["cpu_baseline"] = {
            "value": b_cpu / sec / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "host_cpus": os.cpu_count(),
            "sample": f"the same workload (pubmed-shaped x{args.replicas} replicas, N={dims['N']}, same F sweep), 2 timed passes "
                      "after 1 warm-up of the pure-torch restatement of model/pygnn/hgnn.py:30-37 on the host cores"}
        del step
    del Xs, Ys
    torch.cuda.empty_cache()
    if not args.no_extras:
        line["epoch"] = epoch_block(dev, 0, 1, args)
        blk, _ = c5_block(dev, 0, 1, args, barrier)
        line["c5"] = blk
    print(json.dumps(line), flush=True)


def run_multi(args, dev, rank, world, barrier):
    from hypergef_b200 import ops
    blk, keep = c5_block(dev, rank, world, args, barrier)
    hg, plan, W, N, Fl, full_bytes = (keep[k] for k in ("hg", "plan", "W", "N", "Fl", "full_bytes"))
    best = "column_sharded"
    if "partitioned" in blk and blk["partitioned"]["ms"] < blk["column_sharded"]["ms"]:
        best = "partitioned"
    ms_step = blk[best]["ms"]
    value = full_bytes / ms_step / 1e6
    # a short clocked re-run of the column-sharded call (clocks DURING a timed region) + the launch count
    Xl = torch.randn(N, Fl, device=dev)
    Yl = torch.empty_like(Xl)
    l0 = plan.kernels_launched()
    with ClockSampler(dev.index) as clocks:
        ms_c = timed_loop(lambda: ops.aggregate(plan, Xl, s1=hg.degE, s2=W, a_out=hg.degV, out=Yl), max(3, min(args.steps, 10)), 3, barrier)
        clocks.sample()
    launches = (plan.kernels_launched() - l0)
    # ---- end to end: every rank's column block from pinned host memory and back
    hx, hy = torch.empty(N, Fl).pin_memory(), torch.empty(N, Fl).pin_memory()
    hx.normal_(generator=torch.Generator().manual_seed(7 + rank))
    pipe = ops.HostPipeline(plan)

    def e2e_step():
        pipe.submit(hx, hy, s1=hg.degE, s2=W, a_out=hg.degV)
        pipe.wait()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_ms = max_over_ranks(timed_loop(e2e_step, e2e_steps, 1, barrier), dev, world)
    del hx, hy, pipe
    probe = pcie_probe(dev, world)
    del Xl, Yl
    torch.cuda.empty_cache()
    epoch = epoch_block(dev, rank, world, args) if not args.no_extras else None
    if rank != 0:
        return
    peak, peak_src = peaks()
    cfg = dict(c5_config(args), N=blk["N"], E=blk["E"], nnz=blk["nnz"],
               parallelism=(f"feature-column sharding x{world} (no data-path collective)" if best == "column_sharded" else
                            f"vertex / hyperedge partition x{world} with NCCL all_to_all boundary exchange"))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": max(3, min(args.steps, 10)), "warmup": 3,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": value, "peak": peak * world, "unit": "GB/s", "frac": value / (peak * world),
                         "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "kernel": "stream form (stage A + stage B) on every rank's share", "algorithmic_bytes_per_step": full_bytes},
            "c5": blk,
            "e2e": {"value": full_bytes / e2e_ms / 1e6, "unit": UNIT, "h2d_bytes_per_step": 4 * N * Fl * world,
                    "d2h_bytes_per_step": 4 * N * Fl * world, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "path": "column-sharded: every rank stages its own column block through pinned host memory",
                    "pcie_probe_all_ranks_at_once": probe},
            "gpu_launches": launches, "clocks": clocks.summary(), "column_sharded_ms_recheck": ms_c}
    if epoch is not None:
        line["epoch"] = epoch
    print(json.dumps(line), flush=True)


def same_gpu_baselines(hg, plan, Xs, Ys, W, features, bytes_f, peak):
    """Reported next to the headline (north star): cuSPARSE's two unfused SpMM calls on the same B200
    (torch.sparse CSR @ dense = cusparseSpMM; H^T then H, scales folded into the CSR values, as
    include/spmm/spmm.cuh:22-77,701-709), the reference's own fused kernel compiled for sm_100a (edge_based_full,
    include/hgnnAgg.cuh:98-131; un-scaled operator, zero-fill + kernel as its extension runs it; when oracle/_ref is
    built), and this repo's own two-pass form (cudaMemset + segment kernel)."""
    from hypergef_b200 import _native, ops
    N, M = hg.num_nodes, hg.num_edges
    degE, degV = hg.degE.reshape(-1), hg.degV.reshape(-1)
    rows_t = torch.repeat_interleave(torch.arange(M, device=W.device), (hg.H_T_csrptr[1:] - hg.H_T_csrptr[:-1]).long())
    HT = torch.sparse_csr_tensor(hg.H_T_csrptr, hg.H_T_colind, (degE * W)[rows_t], size=(M, N))
    rows = torch.repeat_interleave(torch.arange(N, device=W.device), (hg.H_csrptr[1:] - hg.H_csrptr[:-1]).long())
    H = torch.sparse_csr_tensor(hg.H_csrptr, hg.H_colind, degV[rows], size=(N, M))
    ref = reference_kernels()

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e3
    out = []
    for F in features:
        X = Xs[F]
        us_cs = timed(lambda: torch.sparse.mm(H, torch.sparse.mm(HT, X)))
        us_2p = timed(lambda: ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, out=Ys[F], flags=_native.HG_TWO_PASS))
        want = torch.sparse.mm(H, torch.sparse.mm(HT, X))
        ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, out=Ys[F])
        err = ((Ys[F] - want).abs().max() / want.abs().max()).item()
        row = {"F": F, "cusparse_2xspmm_us": us_cs, "cusparse_algorithmic_GBps": bytes_f[F] / us_cs / 1e3,
               "two_pass_us": us_2p, "two_pass_algorithmic_GBps": bytes_f[F] / us_2p / 1e3,
               "max_rel_diff_fused_vs_cusparse": err}
        del want
        if ref is not None:
            try:
                yr, us_ref = ref(0, hg.ngs, M, hg.group_key, hg.group_start, hg.group_end, hg.H_T_colind, X, out=Ys[F], iters=5)
                ours_u = ops.aggregate(plan, X)
                row["reference_kernel_us"] = us_ref
                row["reference_kernel_algorithmic_GBps"] = bytes_f[F] / us_ref / 1e3
                row["max_rel_diff_vs_reference_kernel"] = ((ours_u - yr).abs().max() / yr.abs().max()).item()
                del ours_u
            except Exception as exc:
                row["reference_kernel_error"] = repr(exc)[:200]
        out.append(row)
    return out


def layer_block(hg, W):
    """one HGNN layer  degV H (degE W) H^T (X Theta)  on the bench graph, forward, fp32, with the
    projection on the N vertex rows (the reference's order, model/ugsys/hgnn.py:22-23: Linear then HGNNAggr), on the E
    hyperedge rows between the two stages (hg_plan_edge_reduce -> GEMM -> hg_plan_edge_scatter), or after the
    aggregation; torch.matmul fp32 (no TF32) for the GEMM in every arm."""
    from hypergef_b200 import ops
    N, M = hg.num_nodes, hg.num_edges
    out = []
    for f_in, f_out in ((128, 128), (256, 256), (512, 256)):
        X = torch.randn(N, f_in, device=W.device)
        T = torch.randn(f_in, f_out, device=W.device) / f_in ** 0.5
        row = {"F_in": f_in, "F_out": f_out, "chosen": ops.projection_order(N, M, f_in, f_out)}
        with torch.no_grad():
            for order in ("vertex", "edge", "after"):
                fn = lambda: ops.projected_aggregate(hg, X, T, hg.degE, hg.degV, W, order=order)
                for _ in range(3):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    fn()
                b.record()
                torch.cuda.synchronize()
                row[order + "_ms"] = a.elapsed_time(b) / 10
        # forward + backward (gradients of X and Theta), the reference's order against the chosen one
        Xg, Tg = X.clone().requires_grad_(True), T.clone().requires_grad_(True)
        G = torch.randn(N, f_out, device=W.device)
        for order in dict.fromkeys(("vertex", row["chosen"])):
            def fb():
                Xg.grad = Tg.grad = None
                ops.projected_aggregate(hg, Xg, Tg, hg.degE, hg.degV, W, order=order).backward(G)
            for _ in range(3):
                fb()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fb()
            b.record()
            torch.cuda.synchronize()
            row[order + "_fwd_bwd_ms"] = a.elapsed_time(b) / 5
        del Xg, Tg, G
        out.append(row)
        del X, T
    return out


def reference_kernels():
    """The reference's own GPU kernels, compiled in place into oracle/_ref/libhgref.so (a BASELINE that is timed,
    never a path of the product); None when that library was not built."""
    try:
        from oracle import oracle as orc
        return orc.ref_lab_gpu if orc.ref_available() else None
    except Exception:
        return None


def literal_shapes(dev):
    """BASELINE configs at their LITERAL sizes (L2-resident on a B200: latency, not roofline): ours (library's own
    choice of form, warm, back to back) vs the reference's kernel on the same GPU, un-scaled operator."""
    import hypergef_b200 as hgef
    from hypergef_b200 import ops, synth
    ref = reference_kernels()
    out = []
    for shape, F in (("cora", 32), ("pubmed", 64), ("pubmed", 128), ("dblp", 128), ("walmart", 32), ("walmart", 128)):
        data = synth.make_shape(shape, seed=0)
        hg = hgef.HyperGraph(data, dev, data.dataset)
        plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, hg.num_nodes, hg.num_edges)
        X = torch.randn(hg.num_nodes, F, device=dev)
        Y = torch.empty_like(X)
        for _ in range(5):
            ops.aggregate(plan, X, out=Y)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            ops.aggregate(plan, X, out=Y)
        b.record()
        torch.cuda.synchronize()
        row = {"shape": shape, "N": hg.num_nodes, "E": hg.num_edges, "nnz": int(hg.H_T_colind.numel()), "F": F,
               "ours_us": a.elapsed_time(b) / 50 * 1e3}
        if ref is not None:
            try:
                yr, us = ref(0, hg.ngs, hg.num_edges, hg.group_key, hg.group_start, hg.group_end, hg.H_T_colind, X, iters=50)
                row["reference_kernel_us"] = us
                row["max_rel_diff"] = ((Y - yr).abs().max() / yr.abs().max()).item()
            except Exception as exc:
                row["reference_kernel_error"] = repr(exc)[:200]
        out.append(row)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--features", type=lambda s: [int(x) for x in s.split(",")], default=list(FEATURES))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--c5-scale", type=float, default=0.2, help="fraction of the C5 shape (50 M x 10 M) used by the multi-GPU runs")
    ap.add_argument("--ref-c5-scale", type=float, default=0.02, help="bounded sample of the C5 graph for the CPU reference arm")
    ap.add_argument("--c5-feature", type=int, default=256)
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--cpu-epochs", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the same-GPU baselines, the epoch block and the C5 block")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; hypergef_b200 has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
