"""The autograd ops of the fused aggregation path, over the C-ABI.

Mirrors the two torch extensions of the reference (argument order and meaning kept):

* ``hgnnaggr(balan_key, balan_row, group_st, group_ed, csrptr_t, indices_t, node_feat, degE, degV, W)``
  -- ``HyperGsys/source/hgnnaggr/hgnnaggr.cc:122-129`` (autograd ``:37-65``)
* ``hgnnaggr_mean / hgnnaggr_max(csrptr_t, indices_t, node_feat, degE, degV, W)`` -- ``:131-144``
* ``unignnaggrdeg(key, row, st, ed, csrptr_t, indices_t, node_feat, degE, degV)`` and
  ``unignnaggr(key, row, st, ed, csrptr_t, indices_t, node_feat)``
  -- ``HyperGsys/source/unignnaggr/unignnaggr.cc:81-96``

Differences, all deliberate:

* backward is the TRUE transpose ``H S H^T diag(degV) dY`` by default; the reference re-runs
  the forward on ``dY`` (``hgnnaggr.cc:58-60``), which is only the gradient when ``degV`` is
  uniform.  ``set_backward_mode("reference")`` reproduces the reference.
* ``W`` gets a gradient when it requires one (the reference returns none, ``hgnnaggr.cc:62-63``).
* bad dtypes / devices / shapes raise ``TypeError`` / ``ValueError`` (the reference aborts
  the process through C ``assert``, ``hgnnaggr_cuda.cu:8-12``).
* any feature length works (the reference needs ``F < 32`` or ``F % 32 == 0``).
"""
from __future__ import annotations

import ctypes as C
import threading
from collections import OrderedDict

import os

import torch

from . import _native

__all__ = [
    "hgnnaggr", "hgnnaggr_mean", "hgnnaggr_max", "unignnaggrdeg", "unignnaggr",
    "set_backward_mode", "get_backward_mode", "aggregate", "aggregate_host", "HostPipeline", "Plan", "get_plan", "clear_plan_cache",
    "launch_count", "tune", "edge_reduce", "edge_scatter", "projected_aggregate", "projection_order",
]

DEFAULT_FLAGS = 0       # OR-ed into every hg_aggr_forward call (tests force one kernel form with it)
_BACKWARD_MODE = "transpose"
_LAUNCHES = 0           # C-ABI aggregation calls issued (bench.py reports it)


def set_backward_mode(mode: str) -> None:
    """``"transpose"`` (exact gradient, default) or ``"reference"`` (hgnnaggr.cc:58-60)."""
    global _BACKWARD_MODE
    if mode not in ("transpose", "reference"):
        raise ValueError("backward mode must be 'transpose' or 'reference'")
    _BACKWARD_MODE = mode


def get_backward_mode() -> str:
    return _BACKWARD_MODE


def launch_count() -> int:
    return _LAUNCHES


def tune(**knobs) -> None:
    """Override launch-geometry defaults of the aggregation kernels (``hg_tune_set``); ``None`` restores one.

    Measurement / test hook, e.g. ``tune(ring_kb=96, ring_lag_b=600)``; names are listed in DESIGN.md section 3."""
    for name, value in knobs.items():
        _native.call("hg_tune_set", name.encode(), 0 if value is None else int(value), 1 if value is None else 0)


# ----------------------------------------------------------------------------
# argument checking
# ----------------------------------------------------------------------------
_I32_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()   # (data_ptr, numel, version) -> (source, int32 copy)
_I32_CACHE_SIZE = 64


def _index(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (hypergef_b200 has no CPU path)")
    if t.dtype == torch.int64:
        # int64 index tensors (accepted for large graphs) are converted ONCE: the copy is cached against the
        # source tensor, so the plan cache (keyed on data pointers) hits on the next call instead of
        # rebuilding the plan from a fresh copy every time
        ident = (t.data_ptr(), t.numel(), t._version, str(t.device))
        with _PLANS_LOCK:
            hit = _I32_CACHE.get(ident)
            if hit is not None:
                _I32_CACHE.move_to_end(ident)
                return hit[1]
        c = t.to(torch.int32).contiguous().reshape(-1)
        with _PLANS_LOCK:
            _I32_CACHE[ident] = (t, c)          # holding `t` keeps its data_ptr from being recycled
            while len(_I32_CACHE) > _I32_CACHE_SIZE:
                _I32_CACHE.popitem(last=False)
        return c
    if t.dtype != torch.int32:
        raise TypeError(f"{name} must be int32 (or int64), got {t.dtype}")
    return t.contiguous().reshape(-1)


def _feat(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (hypergef_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{name} must be [rows, F], got shape {tuple(t.shape)}")
    return t.contiguous()


def _scale(t, name, n, device):
    """degE [M,1] / degV [N,1] / W [M]: read as flat float arrays (hgnnaggr_cuda.cu:29,43)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if t.device != device:
        raise ValueError(f"{name} is on {t.device}, features are on {device}")
    if t.numel() != n:
        raise ValueError(f"{name} has {t.numel()} entries, expected {n}")
    return t.detach().contiguous().reshape(-1)


def _ptr(t):
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------
# plans
# ----------------------------------------------------------------------------
class Plan:
    """Owns an ``hgPlan`` and keeps the borrowed index tensors alive."""

    def __init__(self, key, row, st, ed, indices_t, num_nodes, num_edges):
        self.tensors = (key, row, st, ed, indices_t)
        self.device = key.device
        self.num_nodes, self.num_edges = int(num_nodes), int(num_edges)
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device_index = dev
        handle = C.c_void_p()
        _native.call("hg_plan_create", C.byref(handle), self.num_nodes, self.num_edges,
                     indices_t.numel(), key.numel() - 1, row.numel(), key.data_ptr(), row.data_ptr(),
                     st.data_ptr(), ed.data_ptr(), indices_t.data_ptr(), dev,
                     torch.cuda.current_stream(dev).cuda_stream)
        self.handle = handle
        nseg, he, hs, canon = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        _native.call("hg_plan_info", handle, C.byref(nseg), C.byref(he), C.byref(hs), C.byref(canon))
        self.nseg, self.nheavy_edges, self.nheavy_segs = nseg.value, he.value, hs.value
        self.canonical = bool(canon.value)

    def kernels_launched(self) -> int:
        """This library's kernels launched through the plan so far (hg_plan_launches)."""
        n = C.c_int64()
        _native.call("hg_plan_launches", self.handle, C.byref(n))
        return n.value

    def reserve(self, F_max: int) -> None:
        """Size the plan's per-call buffers for feature lengths up to ``F_max`` (hg_plan_reserve): after this no
        aggregation call allocates -- required before capturing calls with a new, wider F into a CUDA graph."""
        _native.call("hg_plan_reserve", self.handle, int(F_max), torch.cuda.current_stream(self.device_index).cuda_stream)

    def debug_words(self):
        """The 8 control words of the last ring-form launch (hg_plan_debug; diagnostic)."""
        out = (C.c_int32 * 8)()
        _native.call("hg_plan_debug", self.handle, out, torch.cuda.current_stream(self.device_index).cuda_stream)
        return list(out)

    def check(self) -> None:
        """Synchronise and raise if any launch issued with this plan faulted (hg_plan_check)."""
        _native.call("hg_plan_check", self.handle, torch.cuda.current_stream(self.device_index).cuda_stream)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _native.lib().hg_plan_destroy(h)
            except Exception:
                pass


_PLANS: "OrderedDict[tuple, Plan]" = OrderedDict()
_PLANS_LOCK = threading.Lock()
_PLAN_CACHE_SIZE = 16


def clear_plan_cache() -> None:
    with _PLANS_LOCK:
        _PLANS.clear()
        _I32_CACHE.clear()


def get_plan(key, row, st, ed, indices_t, num_nodes, num_edges) -> Plan:
    """Plan for this balancer output, cached on the identity of the five index tensors.

    The cache entry holds the tensors, so a pointer cannot be recycled while its plan lives.
    Index arrays are treated as immutable once used (as the reference's HyperGraph does).
    """
    key, row, st, ed, indices_t = (_index(t, n) for t, n in (
        (key, "balan_key"), (row, "balan_row"), (st, "group_st"), (ed, "group_ed"),
        (indices_t, "indices_t")))
    if key.numel() < 2:
        raise ValueError("balan_key needs at least one segment and the sentinel")
    if not (row.numel() == st.numel() == ed.numel()):
        raise ValueError("balan_row, group_st and group_ed must have the same length")
    ident = (key.data_ptr(), row.data_ptr(), st.data_ptr(), ed.data_ptr(), indices_t.data_ptr(),
             key.numel(), row.numel(), indices_t.numel(), int(num_nodes), int(num_edges), str(key.device))
    with _PLANS_LOCK:
        plan = _PLANS.get(ident)
        if plan is not None:
            _PLANS.move_to_end(ident)
            return plan
    plan = Plan(key, row, st, ed, indices_t, num_nodes, num_edges)
    with _PLANS_LOCK:
        _PLANS[ident] = plan
        while len(_PLANS) > _PLAN_CACHE_SIZE:
            _PLANS.popitem(last=False)
    return plan


def aggregate(plan: Plan, X, s1=None, s2=None, a_out=None, a_in=None, out=None, flags=0):
    """``Y = diag(a_out) H diag(s1*s2) H^T diag(a_in) X`` -- one call of ``hg_aggr_forward``."""
    global _LAUNCHES
    X = _feat(X, "node_feat")
    if X.device != plan.device:
        raise ValueError(f"node_feat is on {X.device}, the graph is on {plan.device}")
    N, M, F = plan.num_nodes, plan.num_edges, X.shape[1]
    if X.shape[0] != N:
        raise ValueError(f"node_feat has {X.shape[0]} rows, the graph has {N} vertices")
    s1 = _scale(s1, "degE", M, X.device)
    s2 = _scale(s2, "W", M, X.device)
    a_out = _scale(a_out, "degV", N, X.device)
    a_in = _scale(a_in, "degV", N, X.device)
    if out is None:
        out = torch.empty((N, F), dtype=torch.float32, device=X.device)
    if F == 0 or N == 0:
        return out.zero_() if not (flags & _native.HG_ACCUMULATE) else out
    stream = torch.cuda.current_stream(plan.device_index).cuda_stream
    _native.call("hg_aggr_forward", plan.handle, X.data_ptr(), _ptr(s1), _ptr(s2), _ptr(a_out),
                 _ptr(a_in), out.data_ptr(), F, flags | DEFAULT_FLAGS, stream)
    _LAUNCHES += 1
    return out


def edge_reduce(plan: Plan, X, s1=None, s2=None, a_in=None, out=None):
    """Stage A alone: ``Xe = diag(s1*s2) H^T diag(a_in) X`` -> ``[num_edges, F]`` (``hg_plan_edge_reduce``:
    the plan's balanced stream kernel with the hyperedge features in a caller-owned buffer).  F % 4 == 0."""
    global _LAUNCHES
    X = _feat(X, "node_feat")
    N, M, F = plan.num_nodes, plan.num_edges, X.shape[1]
    if X.device != plan.device or X.shape[0] != N:
        raise ValueError(f"node_feat must be [{N}, F] on {plan.device}, got {tuple(X.shape)} on {X.device}")
    s1, s2 = _scale(s1, "degE", M, X.device), _scale(s2, "W", M, X.device)
    a_in = _scale(a_in, "degV", N, X.device)
    if out is None:
        out = torch.empty((M, F), dtype=torch.float32, device=X.device)
    stream = torch.cuda.current_stream(plan.device_index).cuda_stream
    _native.call("hg_plan_edge_reduce", plan.handle, X.data_ptr(), _ptr(s1), _ptr(s2), _ptr(a_in), out.data_ptr(), F, stream)
    _LAUNCHES += 1
    return out


def edge_scatter(plan: Plan, Xe, a_out=None, out=None):
    """Stage B alone: ``Y = diag(a_out) H Xe`` -> ``[num_nodes, F]`` (``hg_plan_edge_scatter``; every row of Y
    is written exactly once).  F % 4 == 0."""
    global _LAUNCHES
    Xe = _feat(Xe, "edge_feat")
    N, M, F = plan.num_nodes, plan.num_edges, Xe.shape[1]
    if Xe.device != plan.device or Xe.shape[0] != M:
        raise ValueError(f"edge_feat must be [{M}, F] on {plan.device}, got {tuple(Xe.shape)} on {Xe.device}")
    a_out = _scale(a_out, "degV", N, Xe.device)
    if out is None:
        out = torch.empty((N, F), dtype=torch.float32, device=Xe.device)
    stream = torch.cuda.current_stream(plan.device_index).cuda_stream
    _native.call("hg_plan_edge_scatter", plan.handle, Xe.data_ptr(), _ptr(a_out), out.data_ptr(), F, stream)
    _LAUNCHES += 1
    return out


# Rough per-unit costs on a B200 for choosing where a layer's projection goes (measured: stream stages move
# ~4.3 TB/s of their own traffic; torch's fp32 matmul ~60 TFLOP/s).  Only the ORDER of the three candidates matters.
_STAGE_BYTES_PER_S = 4.3e12
_GEMM_FLOPS_PER_S = 60e12


def projection_order(num_nodes, num_edges, f_in, f_out) -> str:
    """Where ``Theta`` goes in ``Y = degV H (degE W) H^T (X Theta)``: ``'vertex'`` projects the N vertex rows first
    (the reference: model/ugsys/hgnn.py:22-23), ``'edge'`` projects the E hyperedge rows between the two stages
    (``H^T (X Theta) = (H^T X) Theta``: fewer rows whenever E < N, at the price of running stage A at ``f_in``
    columns), ``'after'`` projects the aggregated vertex rows.  Picks the cheapest by a byte / flop count."""
    N, M = float(num_nodes), float(num_edges)
    stage = lambda f: 4.0 * f * (N + M) / _STAGE_BYTES_PER_S      # one stage at f columns: reads + writes
    gemm = lambda rows: 2.0 * rows * f_in * f_out / _GEMM_FLOPS_PER_S
    cost = {"vertex": gemm(N) + 2 * stage(f_out), "edge": stage(f_in) + gemm(M) + stage(f_out),
            "after": 2 * stage(f_in) + gemm(N)}
    if f_in % 4 or f_out % 4:
        cost.pop("edge")           # the stage entry points take rows of whole 128-bit vectors
    return min(cost, key=cost.get)


class _ProjectedAggr(torch.autograd.Function):
    """``Y = degV H [(degE W H^T X) Theta]``: stage A at F_in, one GEMM over the E hyperedge rows, stage B at F_out
    (SURVEY.md 8(f) N1).  Backward is the exact transpose: ``dZ = H^T degV G``, ``dTheta = Xe^T dZ``,
    ``dX = H (degE W) (dZ Theta^T)`` -- the same two stage kernels and two GEMMs over E rows."""

    @staticmethod
    def forward(ctx, plan, X, theta, degE, degV, W):
        Xe = edge_reduce(plan, X, s1=degE, s2=W)
        Y = edge_scatter(plan, Xe @ theta, a_out=degV)
        ctx.plan, ctx.scales = plan, (degE, degV, W)
        ctx.save_for_backward(Xe, theta)
        return Y

    @staticmethod
    def backward(ctx, G):
        plan = ctx.plan
        degE, degV, W = ctx.scales
        Xe, theta = ctx.saved_tensors
        dZ = edge_reduce(plan, G.contiguous(), a_in=degV)
        grad_theta = Xe.t() @ dZ if ctx.needs_input_grad[2] else None
        grad_x = None
        if ctx.needs_input_grad[1]:
            dXe = dZ @ theta.t()
            if degE is not None:
                dXe.mul_(degE.detach().reshape(-1, 1))
            if W is not None:
                dXe.mul_(W.detach().reshape(-1, 1))
            grad_x = edge_scatter(plan, dXe)
        return None, grad_x, grad_theta, None, None, None


def projected_aggregate(hyperg, X, theta, degE=None, degV=None, W=None, order="auto"):
    """One HGNN layer ``degV H (degE W) H^T (X theta)`` with ``theta`` ``[F_in, F_out]`` applied where it is
    cheapest (``projection_order``; ``order`` forces ``'vertex'`` / ``'edge'`` / ``'after'``).  ``'vertex'`` is
    exactly what the reference's conv module does (``nn.Linear`` then ``HGNNAggr``); the other two give the same
    result up to fp32 rounding of the re-associated sums.  Falls back to ``'vertex'`` when ``W`` needs a gradient
    or the reference's backward is selected (those paths exist only for the fused op)."""
    N = X.shape[0]
    M = degE.numel() if degE is not None else hyperg.H_T_csrptr.numel() - 1
    f_in, f_out = theta.shape
    if order == "auto":
        order = projection_order(N, M, f_in, f_out)
    if order not in ("vertex", "edge", "after"):
        raise ValueError(f"order must be 'auto', 'vertex', 'edge' or 'after', got {order!r}")
    if order == "edge" and ((W is not None and W.requires_grad) or _BACKWARD_MODE != "transpose"):
        order = "vertex"
    args = (hyperg.group_key, hyperg.group_row, hyperg.group_start, hyperg.group_end, hyperg.H_T_csrptr, hyperg.H_T_colind)
    if order == "vertex":
        return _FusedAggr.apply(*args, X @ theta, degE, degV, W, N)
    if order == "after":
        return _FusedAggr.apply(*args, X, degE, degV, W, N) @ theta
    if f_in % 4 or f_out % 4:
        raise ValueError(f"order='edge' needs feature lengths that are multiples of 4, got {f_in} -> {f_out}")
    plan = get_plan(*args[:4], args[5], N, M)
    return _ProjectedAggr.apply(plan, _feat(X, "node_feat"), theta, degE, degV, W)


class HostPipeline:
    """Aggregation for HOST feature matrices (numpy / CPU torch callers), copies overlapped.

    Three streams: host->device copies, the aggregation launches, device->host copies.  A matrix is
    staged in COLUMN SLABS (the aggregation never mixes feature columns): every slab is its own
    upload -> launch -> download through a small ring of preallocated device buffers, so the upload of
    slab k+1 and the download of slab k-1 run while slab k computes (PCIe is full duplex) -- inside one
    wide matrix as well as across calls -- and no allocation happens on the hot path.  All launches
    share ONE compute stream, so the plan's per-call scratch is never used by two launches at once.
    ``wait()`` blocks until every submitted result is in its host buffer.  Pinned host buffers make
    the copies real DMA.
    """

    def __init__(self, plan: Plan, col_slab: int = 128, depth: int = 4):
        self.plan = plan
        # columns per staged slab (0 = whole matrices, device buffers from the caching allocator);
        # HGEF_COL_SLAB overrides the default for experiments
        self.col_slab = int(os.environ.get("HGEF_COL_SLAB", col_slab))
        self.depth = max(2, int(depth))
        with torch.cuda.device(plan.device_index):
            self.h2d, self.compute, self.d2h = (torch.cuda.Stream() for _ in range(3))
        self._slots, self._next = [], 0
        self._keep = []

    def _slot(self, width):
        """Next staging slot (device X / Y buffers of ``num_nodes x col_slab`` floats + a 'free again' event)."""
        if not self._slots:
            n = self.plan.num_nodes * self.col_slab
            for _ in range(self.depth):
                self._slots.append({"X": torch.empty(n, dtype=torch.float32, device=self.plan.device),
                                    "Y": torch.empty(n, dtype=torch.float32, device=self.plan.device), "free": None})
        slot = self._slots[self._next]
        self._next = (self._next + 1) % self.depth
        n = self.plan.num_nodes * width
        return slot, slot["X"][:n].view(self.plan.num_nodes, width), slot["Y"][:n].view(self.plan.num_nodes, width)

    def submit(self, X_host, out_host=None, s1=None, s2=None, a_out=None, a_in=None):
        plan = self.plan
        if not isinstance(X_host, torch.Tensor):
            X_host = torch.as_tensor(X_host)
        if X_host.is_cuda or X_host.dtype != torch.float32 or X_host.dim() != 2:
            raise TypeError("aggregate_host needs a 2-D float32 CPU tensor")
        X_host = X_host.contiguous()
        if out_host is None:
            out_host = torch.empty_like(X_host, pin_memory=X_host.is_pinned())
        if out_host.shape != X_host.shape or out_host.dtype != torch.float32 or out_host.is_cuda \
                or not out_host.is_contiguous():
            raise ValueError("out_host must be a contiguous float32 CPU tensor of X_host's shape")
        N, F = X_host.shape
        if N != plan.num_nodes:
            raise ValueError(f"X_host has {N} rows, the graph has {plan.num_nodes} vertices")
        dev = plan.device_index
        with torch.cuda.device(dev):
            # device inputs (the scale vectors) may have been produced on the caller's stream just now: the
            # private streams start after it, and the caching allocator is told who else uses those tensors
            cur = torch.cuda.current_stream(dev)
            self.h2d.wait_stream(cur)
            self.compute.wait_stream(cur)
            for t in (s1, s2, a_out, a_in):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(self.compute)
            if self.col_slab <= 0:          # whole matrix, buffers from the caching allocator
                with torch.cuda.stream(self.h2d):
                    Xd = X_host.to(plan.device, non_blocking=True)
                    up = torch.cuda.Event()
                    up.record()
                with torch.cuda.stream(self.compute):
                    self.compute.wait_event(up)
                    Yd = aggregate(plan, Xd, s1=s1, s2=s2, a_out=a_out, a_in=a_in)
                    Xd.record_stream(self.compute)
                    done = torch.cuda.Event()
                    done.record()
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(done)
                    out_host.copy_(Yd, non_blocking=True)
                    Yd.record_stream(self.d2h)
            else:
                for c0 in range(0, F, self.col_slab):
                    w = min(self.col_slab, F - c0)
                    slot, Xd, Yd = self._slot(w)
                    with torch.cuda.stream(self.h2d):
                        if slot["free"] is not None:
                            self.h2d.wait_event(slot["free"])      # its previous download has finished
                        _native.call("hg_copy_columns", Xd.data_ptr(), X_host.data_ptr(), N, w, F, c0, 1, dev,
                                     self.h2d.cuda_stream)
                        up = torch.cuda.Event()
                        up.record()
                    with torch.cuda.stream(self.compute):
                        self.compute.wait_event(up)
                        aggregate(plan, Xd, s1=s1, s2=s2, a_out=a_out, a_in=a_in, out=Yd)
                        done = torch.cuda.Event()
                        done.record()
                    with torch.cuda.stream(self.d2h):
                        self.d2h.wait_event(done)
                        _native.call("hg_copy_columns", out_host.data_ptr(), Yd.data_ptr(), N, w, F, c0, 0, dev,
                                     self.d2h.cuda_stream)
                        slot["free"] = torch.cuda.Event()
                        slot["free"].record()
        self._keep.append((X_host, out_host))
        return out_host

    def wait(self):
        self.d2h.synchronize()
        self._keep.clear()


def aggregate_host(plan: Plan, X_host, out_host=None, s1=None, s2=None, a_out=None, a_in=None):
    """One blocking host-buffer aggregation: ``X_host`` [N,F] fp32 -> device -> ``hg_aggr_forward``
    -> ``out_host``.  For several matrices use :class:`HostPipeline` so the copies overlap."""
    pipe = getattr(plan, "_host_pipe", None)
    if pipe is None:
        pipe = plan._host_pipe = HostPipeline(plan, depth=2)
    out = pipe.submit(X_host, out_host, s1=s1, s2=s2, a_out=a_out, a_in=a_in)
    pipe.wait()
    return out


def _weight_grad(csrptr_t, indices_t, X, G, s1, a_out, a_in, M):
    dW = torch.empty(M, dtype=torch.float32, device=X.device)
    dev = X.device.index if X.device.index is not None else torch.cuda.current_device()
    _native.call("hg_weight_grad", M, csrptr_t.data_ptr(), indices_t.data_ptr(), X.data_ptr(),
                 G.data_ptr(), _ptr(s1), _ptr(a_out), _ptr(a_in), dW.data_ptr(), X.shape[1], dev,
                 torch.cuda.current_stream(dev).cuda_stream)
    return dW


# ----------------------------------------------------------------------------
# autograd functions
# ----------------------------------------------------------------------------
class _FusedAggr(torch.autograd.Function):
    """HGNNAggr / UniGNNAggrDeg / UniGNNAggr in one Function (scales optional)."""

    @staticmethod
    def forward(ctx, key, row, st, ed, csrptr_t, indices_t, node_feat, degE, degV, W, num_nodes):
        node_feat = _feat(node_feat, "node_feat")
        M = degE.numel() if degE is not None else (_index(csrptr_t, "csrptr_t").numel() - 1)
        plan = get_plan(key, row, st, ed, indices_t, num_nodes, M)
        out = aggregate(plan, node_feat, s1=degE, s2=W, a_out=degV)
        ctx.plan, ctx.scales = plan, (degE, degV, W)
        ctx.csr = (csrptr_t, indices_t)
        ctx.w_needs_grad = W is not None and ctx.needs_input_grad[9]
        if ctx.w_needs_grad:
            ctx.save_for_backward(node_feat)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        degE, degV, W = ctx.scales
        grad_out = grad_out.contiguous()
        grad_x = grad_w = None
        if ctx.needs_input_grad[6]:
            if _BACKWARD_MODE == "reference":       # hgnnaggr.cc:58-60: forward applied to grad_out
                grad_x = aggregate(ctx.plan, grad_out, s1=degE, s2=W, a_out=degV)
            else:                                   # true transpose: degV moves to the gather side
                grad_x = aggregate(ctx.plan, grad_out, s1=degE, s2=W, a_in=degV)
        if ctx.w_needs_grad:
            (x,) = ctx.saved_tensors
            csrptr_t, indices_t = (_index(t, n) for t, n in zip(ctx.csr, ("csrptr_t", "indices_t")))
            s1 = _scale(degE, "degE", ctx.plan.num_edges, x.device)
            a_out = _scale(degV, "degV", ctx.plan.num_nodes, x.device)
            grad_w = _weight_grad(csrptr_t, indices_t, x, grad_out, s1, a_out, None,
                                  ctx.plan.num_edges).reshape(W.shape)
        return (None,) * 6 + (grad_x, None, None, grad_w, None)


def hgnnaggr(balan_key, balan_row, group_st, group_ed, csrptr_t, indices_t, node_feat, degE, degV, W):
    """``Y = degV . H . (degE*W) . H^T . X`` (hgnnaggr.cc:122-129).  N is ``degV.size(0)``
    as in the reference launcher (hgnnaggr_cuda.cu:366)."""
    return _FusedAggr.apply(balan_key, balan_row, group_st, group_ed, csrptr_t, indices_t, node_feat,
                            degE, degV, W, degV.shape[0])


def unignnaggrdeg(balan_key, balan_row, group_st, group_ed, csrptr_t, indices_t, node_feat, degE, degV):
    """``Y = degV . H . degE . H^T . X`` (unignnaggr.cc:81-88).  Uses ``degV[v]`` -- the
    reference's non-smem kernels index degV by nnz position (unignnaggr_cuda.cu:41,212)."""
    return _FusedAggr.apply(balan_key, balan_row, group_st, group_ed, csrptr_t, indices_t, node_feat,
                            degE, degV, None, degV.shape[0])


def unignnaggr(balan_key, balan_row, group_st, group_ed, csrptr_t, indices_t, node_feat):
    """``Y = H . H^T . X`` (unignnaggr.cc:90-96).  N is ``in_feat.size(0)`` (unignnaggr_cuda.cu:460)."""
    return _FusedAggr.apply(balan_key, balan_row, group_st, group_ed, csrptr_t, indices_t, node_feat,
                            None, None, None, node_feat.shape[0])


# ---- first-stage mean / max (hgnnaggr.cc:67-120) ----------------------------------------
def _csr_args(csrptr_t, indices_t, node_feat, degE, degV, W):
    csrptr_t, indices_t = _index(csrptr_t, "csrptr_t"), _index(indices_t, "indices_t")
    X = _feat(node_feat, "node_feat")
    M, N = csrptr_t.numel() - 1, X.shape[0]
    if degV is not None and degV.numel() != N:
        raise ValueError(f"degV has {degV.numel()} entries, node_feat has {N} rows")
    return (csrptr_t, indices_t, X, _scale(degE, "degE", M, X.device), _scale(degV, "degV", N, X.device),
            _scale(W, "W", M, X.device), N, M)


def _dev_stream(t):
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return dev, torch.cuda.current_stream(dev).cuda_stream


MEAN_NGS = 64          # balancer segment size of the plan built for the (un-balanced) mean / max entry points
MEAN_BALANCED = True   # False: the reference's own one-warp-per-hyperedge scheme (hg_aggr_mean), kept for A/B


def _csr_plan(csrptr_t, indices_t, N, M) -> Plan:
    """Plan for an UN-balanced ``H^T`` CSR: the reference's mean / max ops take no balancer arrays
    (hgnnaggr.cc:131-144) and walk a whole hyperedge with one warp; here the device balancer cuts the rows into
    segments first, so a 12 345-member hyperedge is shared by many warps like in the sum op.  Cached on the CSR."""
    from .balancer import balance_schedule
    ident = ("csr", csrptr_t.data_ptr(), indices_t.data_ptr(), csrptr_t.numel(), indices_t.numel(), N, M, str(csrptr_t.device))
    with _PLANS_LOCK:
        plan = _PLANS.get(ident)
        if plan is not None:
            _PLANS.move_to_end(ident)
            return plan
    bs = balance_schedule(MEAN_NGS, csrptr_t)
    plan = Plan(bs.balan_key, bs.balan_row, bs.group_st, bs.group_ed, indices_t, N, M)
    plan.tensors = plan.tensors + (csrptr_t,)
    with _PLANS_LOCK:
        _PLANS[ident] = plan
        while len(_PLANS) > _PLAN_CACHE_SIZE:
            _PLANS.popitem(last=False)
    return plan


def _mean_scale(csrptr_t, s1):
    """degE[e] / |e|: the mean over the members folded into the hyperedge scale (hgnnaggr_cuda.cu:107)."""
    deg = (csrptr_t[1:] - csrptr_t[:-1]).to(torch.float32).clamp_(min=1.0)
    return (1.0 / deg) if s1 is None else (s1 / deg)


class _MeanF1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, csrptr_t, indices_t, node_feat, degE, degV, W):
        csrptr_t, indices_t, X, s1, a_out, s2, N, M = _csr_args(csrptr_t, indices_t, node_feat, degE, degV, W)
        ctx.args = (csrptr_t, indices_t, s1, a_out, s2, N, M)
        return _MeanF1._apply(ctx.args, X)

    @staticmethod
    def _apply(args, X):
        csrptr_t, indices_t, s1, a_out, s2, N, M = args
        if MEAN_BALANCED and M > 0 and indices_t.numel() > 0:
            # mean = sum with degE / |e| as the hyperedge scale: the balanced sum kernels do the work
            return aggregate(_csr_plan(csrptr_t, indices_t, N, M), X, s1=_mean_scale(csrptr_t, s1), s2=s2, a_out=a_out)
        out = torch.empty_like(X)
        dev, stream = _dev_stream(X)
        _native.call("hg_aggr_mean", N, M, csrptr_t.data_ptr(), indices_t.data_ptr(), X.data_ptr(),
                     _ptr(s1), _ptr(s2), _ptr(a_out), out.data_ptr(), X.shape[1], 0, dev, stream)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        # hgnnaggr_cuda.cu:115-142: the same mean operator applied to grad_out
        return None, None, _MeanF1._apply(ctx.args, grad_out.contiguous()), None, None, None


class _MaxF1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, csrptr_t, indices_t, node_feat, degE, degV, W):
        csrptr_t, indices_t, X, s1, a_out, s2, N, M = _csr_args(csrptr_t, indices_t, node_feat, degE, degV, W)
        out = torch.empty_like(X)
        record = torch.empty((M, X.shape[1]), dtype=torch.int32, device=X.device)
        dev, stream = _dev_stream(X)
        if MEAN_BALANCED and M > 0 and indices_t.numel() > 0:
            plan = _csr_plan(csrptr_t, indices_t, N, M)
            _native.call("hg_plan_max_forward", plan.handle, csrptr_t.data_ptr(), X.data_ptr(), _ptr(s1), _ptr(s2),
                         _ptr(a_out), out.data_ptr(), record.data_ptr(), X.shape[1], stream)
        else:
            _native.call("hg_aggr_max_forward", N, M, csrptr_t.data_ptr(), indices_t.data_ptr(), X.data_ptr(),
                         _ptr(s1), _ptr(s2), _ptr(a_out), out.data_ptr(), record.data_ptr(), X.shape[1], 0,
                         dev, stream)
        ctx.args = (csrptr_t, indices_t, s1, a_out, s2, N, M, record)
        ctx.mark_non_differentiable(record)
        return out, record

    @staticmethod
    def backward(ctx, grad_out, _grad_record):
        csrptr_t, indices_t, s1, a_out, s2, N, M, record = ctx.args
        G = grad_out.contiguous()
        dX = torch.empty_like(G)
        dev, stream = _dev_stream(G)
        if MEAN_BALANCED and M > 0 and indices_t.numel() > 0:
            plan = _csr_plan(csrptr_t, indices_t, N, M)
            _native.call("hg_plan_max_backward", plan.handle, csrptr_t.data_ptr(), G.data_ptr(), _ptr(s1), _ptr(s2),
                         _ptr(a_out), record.data_ptr(), dX.data_ptr(), G.shape[1], stream)
        else:
            _native.call("hg_aggr_max_backward", N, M, csrptr_t.data_ptr(), indices_t.data_ptr(), G.data_ptr(),
                         _ptr(s1), _ptr(s2), _ptr(a_out), record.data_ptr(), dX.data_ptr(), G.shape[1], 0,
                         dev, stream)
        return None, None, dX, None, None, None


def hgnnaggr_mean(csrptr_t, indices_t, node_feat, degE, degV, W):
    """First-stage MEAN (hgnnaggr.cc:131-136)."""
    return _MeanF1.apply(csrptr_t, indices_t, node_feat, degE, degV, W)


def hgnnaggr_max(csrptr_t, indices_t, node_feat, degE, degV, W):
    """First-stage MAX; returns ``[out, record_table]`` (hgnnaggr.cc:138-144)."""
    return list(_MaxF1.apply(csrptr_t, indices_t, node_feat, degE, degV, W))
