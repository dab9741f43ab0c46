#!/usr/bin/env python
"""bench.py -- fused hypergraph aggregation throughput (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, C-ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (host cores)

Workload (``config.workload``): the standalone fused-aggregation sweep, feature length
32..512, on a synthetic Pubmed-shaped hypergraph (19717 vertices / 7963 hyperedges per replica,
ngs=40) stacked block-diagonally ``--replicas`` times so that X and Y are each larger than the
126 MB L2 at every F (literal Pubmed is L2-resident on a B200).  One STEP = one
pass of the hot path over the batch: the HGNN aggregation ``Y = degV.H.(degE*W).H^T.X`` once per
feature length of the sweep.  ``value`` = algorithmic GB/s = sum_F B_alg(F) / step time with
    B_alg(F) = 8*F*N + 4*Z + 12*E + 4*N + 4   bytes  (every array once)
and inputs resident in HBM.  ``e2e`` = the same metric through the host-buffer API (pinned
host X -> device -> aggregation -> host Y inside the timed region).  N > 1: feature-column
sharding (each rank owns its own column block of the same graph; no data-path collective) =
weak scaling.
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


# ----------------------------------------------------------------------------- reference arm
def cpu_conv_workload(replicas, features, seed=0):
    """Graph + features for the reference's PyG-equivalent CPU conv (model/pygnn/hgnn.py:30-37)."""
    from hypergef_b200 import synth
    from oracle import oracle as orc
    data = synth.make_shape("pubmed", replicas=replicas, seed=seed)
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
    if rank != 0:
        return
    features = tuple(args.features)
    reps = args.ref_replicas
    step, bytes_step, dims = cpu_conv_workload(reps, features)
    cores = torch.get_num_threads()
    # bound the whole run to a few minutes: shrink K if one step is slow
    t_probe = time_cpu(step, 1, 1)
    steps = max(1, min(args.steps, int(180.0 / max(t_probe, 1e-3))))
    warmup = min(args.warmup, 3)
    sec = time_cpu(step, steps, warmup)
    val = bytes_step / sec / 1e9
    sample = (f"pubmed-shaped x{reps} replicas (N={dims['N']}, E={dims['E']}, nnz={dims['Z']}), "
              f"F sweep {list(features)}, pure-torch restatement of model/pygnn/hgnn.py:30-37 "
              f"(index_select + index_add_; torch_scatter/PyG are not installed)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, reps),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, replicas):
    return {"workload": f"pubmed-shaped hypergraph x{replicas} block-diagonal replicas, fused HGNN aggregation, "
                        f"F sweep {list(args.features)}",
            "shape": "pubmed (19717 vertices, 7963 hyperedges per replica)", "replicas": replicas,
            "ngs": 40, "features": list(args.features),
            "l2": "X and Y of every launch are each > 126 MB L2 (inputs larger than L2; no flush)",
            "parallelism": f"feature-column sharding x{args.gpus} (no data-path collective)"}


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import hypergef_b200 as hgef
    from hypergef_b200 import ops, synth

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    features = tuple(args.features)
    data = synth.make_shape("pubmed", replicas=args.replicas, seed=0, device=dev)
    hg = hgef.HyperGraph(data, dev, "pubmed")
    N, M, Z = hg.num_nodes, hg.num_edges, int(hg.H_T_colind.numel())
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, N, M)
    W = torch.ones(M, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    Xs = {F: torch.randn(N, F, device=dev, generator=gen) for F in features}
    Ys = {F: torch.empty(N, F, device=dev) for F in features}
    bytes_f = {F: b_alg(N, M, Z, F) for F in features}
    bytes_step = sum(bytes_f.values())

    def call(F):
        ops.aggregate(plan, Xs[F], s1=hg.degE, s2=W, a_out=hg.degV, out=Ys[F])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        for F in features:
            call(F)
    launches0 = plan.kernels_launched()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in features]
          for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
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
    ms_total = t0.elapsed_time(t1)
    per_f_ms = {F: float(np.mean([ev[k][j][0].elapsed_time(ev[k][j][1]) for k in range(args.steps)]))
                for j, F in enumerate(features)}
    ms_t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = ms_t.item() / args.steps
    value = world * bytes_step / (ms_step * 1e-3) / 1e9

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
    e_t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
    e2e_ms = e_t.item() / e2e_steps
    e2e_val = world * bytes_step / (e2e_ms * 1e-3) / 1e9
    io_bytes = sum(4 * F * N for F in features)

    if rank != 0:
        return
    peak, peak_src = peaks()
    kern_ms = sum(per_f_ms.values())
    achieved = bytes_step / (kern_ms * 1e-3) / 1e9
    sweep = [{"F": F, "us": per_f_ms[F] * 1e3, "algorithmic_GBps": bytes_f[F] / (per_f_ms[F] * 1e-3) / 1e9,
              "frac_of_measured_peak": bytes_f[F] / (per_f_ms[F] * 1e-3) / 1e9 / peak} for F in features]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_step")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, args.replicas), N=N, E=M, nnz=Z,
                           heavy_hyperedges=plan.nheavy_edges, segments=plan.nseg),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "hg_aggr_forward per F = stream form: stream_kernel stage A (X -> Xe over balancer segments) + "
                                   "stage B (Xe -> Y over vertices), two launches per call, "
                                   "CUDA events around each C-ABI call",
                         "algorithmic_bytes_per_step": bytes_step, "sweep": sweep},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes,
                    "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": kernels,
            "clocks": clocks.summary()}
    if world == 1 and not args.no_extras:
        line["same_gpu_baselines"] = same_gpu_baselines(hg, plan, Xs, Ys, W, features, bytes_f, peak)
    if world == 1 and not args.no_cpu_baseline:
        step, b_cpu, dims = cpu_conv_workload(args.ref_replicas, features)
        sec = time_cpu(step, 2, 1)
        line["cpu_baseline"] = {
            "value": b_cpu / sec / 1e9, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": f"pubmed-shaped x{args.ref_replicas} replicas (N={dims['N']}), same F sweep, 2 timed passes of "
                      "the pure-torch restatement of model/pygnn/hgnn.py:30-37 on the host cores"}
    print(json.dumps(line), flush=True)


def same_gpu_baselines(hg, plan, Xs, Ys, W, features, bytes_f, peak):
    """Reported next to the headline (north star): cuSPARSE's two unfused SpMM calls on the same B200
    (torch.sparse CSR @ dense = cusparseSpMM; H^T then H, scales folded into the CSR values, as
    include/spmm/spmm.cuh:22-77,701-709) and this repo's own two-pass form (cudaMemset + segment kernel)."""
    from hypergef_b200 import _native, ops
    N, M = hg.num_nodes, hg.num_edges
    degE, degV = hg.degE.reshape(-1), hg.degV.reshape(-1)
    rows_t = torch.repeat_interleave(torch.arange(M, device=W.device), (hg.H_T_csrptr[1:] - hg.H_T_csrptr[:-1]).long())
    HT = torch.sparse_csr_tensor(hg.H_T_csrptr, hg.H_T_colind, (degE * W)[rows_t], size=(M, N))
    rows = torch.repeat_interleave(torch.arange(N, device=W.device), (hg.H_csrptr[1:] - hg.H_csrptr[:-1]).long())
    H = torch.sparse_csr_tensor(hg.H_csrptr, hg.H_colind, degV[rows], size=(N, M))

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
        ref = torch.sparse.mm(H, torch.sparse.mm(HT, X))
        ops.aggregate(plan, X, s1=hg.degE, s2=W, a_out=hg.degV, out=Ys[F])
        err = ((Ys[F] - ref).abs().max() / ref.abs().max()).item()
        out.append({"F": F, "cusparse_2xspmm_us": us_cs, "cusparse_algorithmic_GBps": bytes_f[F] / us_cs / 1e3,
                    "two_pass_us": us_2p, "two_pass_algorithmic_GBps": bytes_f[F] / us_2p / 1e3,
                    "max_rel_diff_fused_vs_cusparse": err})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--ref-replicas", type=int, default=8)
    ap.add_argument("--features", type=lambda s: [int(x) for x in s.split(",")], default=list(FEATURES))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cuSPARSE / two-pass same-GPU baselines")
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
