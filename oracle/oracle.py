"""CPU oracle for the fused hypergraph aggregation path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; nothing under
``hypergef_b200/`` does (tests/test_boundary.py greps for it).

Parity status: PINNED (see hg_oracle.c header).  Three independent statements
of every result are kept so they can be checked against each other and
against the real reference:

* ``c_*``      the plain-C restatement (``hg_oracle.c`` via ctypes);
* ``ref_*``    the reference's own C++ compiled in place (``oracle/_ref``, when built);
* ``py_*`` / ``torch_*``  numpy / scipy / pure-torch restatements of the reference's
               Python (``balancer.py``, ``hypergraph.py``, ``model/pygnn/hgnn.py``).

Citations are relative to the reference checkout.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "libhg_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libhgref.so")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the C oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        _lib = C.CDLL(_ORACLE_SO)
    return _lib


def ref_available() -> bool:
    return os.path.exists(_REF_SO)


_ref = None


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_SO)
    return _ref


def _opt(a):
    """float32 pointer or NULL."""
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))


def _i32(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))


# ----------------------------------------------------------------------------
# balancer
# ----------------------------------------------------------------------------
def py_balancer(ngs: int, csrptr) -> SimpleNamespace:
    """Pure-Python restatement of HyperGsys/balancer.py:15-33 (small inputs only)."""
    csrptr = [int(x) for x in np.asarray(csrptr).reshape(-1)]
    nrow = len(csrptr) - 1
    key, row, st, ed = [], [], [], []
    base = 0
    for rid in range(nrow):
        lb, hb = csrptr[rid], csrptr[rid + 1]
        workload = -(-(hb - lb) // ngs)                      # balancer.py:19
        k = lb
        while k < hb:                                        # :20-23
            key.append(k)
            k += ngs
        for i in range(workload):                            # :24-30
            for j in range(workload):
                st.append(base + j)
                ed.append(base + i)
                row.append(rid)
        base += workload                                     # :31
    if key[-1] != csrptr[-1]:                                # :32-33 (IndexError if key is empty)
        key.append(csrptr[-1])
    return SimpleNamespace(balan_key=key, balan_row=row, group_st=st, group_ed=ed)


def c_balancer(ngs: int, csrptr) -> SimpleNamespace:
    """hg_oracle.c:orc_balance_* (follows balancer.py:15-33 / balancer_kernel.cuh:229-259)."""
    ptr = _i32(csrptr)
    nrow = ptr.size - 1
    nk, ng = C.c_int64(), C.c_int64()
    rc = lib().orc_balance_count(C.c_int64(nrow), ptr.ctypes.data_as(C.c_void_p), C.c_int32(ngs),
                                 C.byref(nk), C.byref(ng))
    if rc:
        raise ValueError("orc_balance_count failed")
    if nk.value == 0:
        raise IndexError("list index out of range")          # balancer.py:32 on an empty key list
    key = np.empty(nk.value, np.int32)
    row, st, ed = (np.empty(ng.value, np.int32) for _ in range(3))
    lib().orc_balance_fill(C.c_int64(nrow), ptr.ctypes.data_as(C.c_void_p), C.c_int32(ngs),
                           key.ctypes.data_as(C.c_void_p), row.ctypes.data_as(C.c_void_p),
                           st.ctypes.data_as(C.c_void_p), ed.ctypes.data_as(C.c_void_p))
    return SimpleNamespace(balan_key=key, balan_row=row, group_st=st, group_ed=ed)


def ref_balancer(ngs: int, csrptr) -> SimpleNamespace:
    """The reference's C++ twin compiled in place (balancer_kernel.cuh:229-259)."""
    ptr = _i32(csrptr)
    nk, ng = C.c_longlong(), C.c_longlong()
    ref_lib().ref_balance_run(C.c_int(ptr.size - 1), C.c_int(ngs), ptr.ctypes.data_as(C.c_void_p),
                              C.byref(nk), C.byref(ng))
    key = np.empty(nk.value, np.int32)
    row, st, ed = (np.empty(ng.value, np.int32) for _ in range(3))
    ref_lib().ref_balance_fetch(key.ctypes.data_as(C.c_void_p), row.ctypes.data_as(C.c_void_p),
                                st.ctypes.data_as(C.c_void_p), ed.ctypes.data_as(C.c_void_p))
    return SimpleNamespace(balan_key=key, balan_row=row, group_st=st, group_ed=ed)


# ----------------------------------------------------------------------------
# incidence / CSR / degrees
# ----------------------------------------------------------------------------
def split_edge_index(edge_index: torch.Tensor, num_nodes: int):
    """hypergraph.py:15-20 -- V, E (hyperedge ids re-based to 0), num_edges, nnz."""
    c_idx = torch.where(edge_index[0] == num_nodes)[0].min()
    V2E = edge_index[:, :c_idx]
    V = V2E[0]
    E = V2E[1] - num_nodes
    return V, E, len(V2E[1].unique()), V2E.shape[1]


def scipy_incidence(V, E, num_nodes: int, num_edges: int):
    """hypergraph.py:22-25 verbatim semantics: scipy coo -> csr, transpose -> csr."""
    import scipy.sparse as sp
    V = np.asarray(V)
    E = np.asarray(E)
    H = sp.coo_matrix((np.ones(V.shape[0]), (V, E)), shape=(num_nodes, num_edges)).tocsr()
    H_T = H.transpose().tocsr()
    return H, H_T


def scipy_degrees(H):
    """hypergraph.py:34-45 -- degV = rowsum^-1/2 (inf -> 1) [N,1], degE = colsum^-1 [M,1]."""
    degV = torch.from_numpy(np.asarray(H.sum(axis=1))).float()
    degE = torch.from_numpy(np.asarray(H.sum(axis=0))).squeeze()
    degE = degE.unsqueeze(1).float() if degE.dim() else degE.reshape(1, 1).float()
    degV = degV.pow(-0.5)
    degE = degE.pow(-1)
    degV[torch.isinf(degV)] = 1
    return degV, degE


def c_csr_from_coo(V, E, num_nodes: int, num_edges: int):
    """hg_oracle.c:orc_csr_from_coo + orc_csr_transpose (scipy semantics)."""
    rows = np.ascontiguousarray(np.asarray(V, dtype=np.int64))
    cols = np.ascontiguousarray(np.asarray(E, dtype=np.int64))
    n = rows.size
    indptr = np.empty(num_nodes + 1, np.int32)
    indices = np.empty(max(n, 1), np.int32)
    data = np.empty(max(n, 1), np.float32)
    out = C.c_int64()
    rc = lib().orc_csr_from_coo(C.c_int64(num_nodes), C.c_int64(num_edges), C.c_int64(n),
                                rows.ctypes.data_as(C.c_void_p), cols.ctypes.data_as(C.c_void_p),
                                indptr.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p),
                                data.ctypes.data_as(C.c_void_p), C.byref(out))
    if rc:
        raise ValueError(f"orc_csr_from_coo failed ({rc})")
    z = out.value
    indices, data = indices[:z].copy(), data[:z].copy()
    t_indptr = np.empty(num_edges + 1, np.int32)
    t_indices = np.empty(max(z, 1), np.int32)
    t_data = np.empty(max(z, 1), np.float32)
    lib().orc_csr_transpose(C.c_int64(num_nodes), C.c_int64(num_edges),
                            indptr.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p),
                            data.ctypes.data_as(C.c_void_p), t_indptr.ctypes.data_as(C.c_void_p),
                            t_indices.ctypes.data_as(C.c_void_p), t_data.ctypes.data_as(C.c_void_p))
    return SimpleNamespace(indptr=indptr, indices=indices, data=data, t_indptr=t_indptr,
                           t_indices=t_indices[:z].copy(), t_data=t_data[:z].copy())


# ----------------------------------------------------------------------------
# aggregation
# ----------------------------------------------------------------------------
def c_aggr_groups(key, row, st, ed, colind, X, s1=None, s2=None, a_out=None, a_in=None,
                  num_nodes=None, f64=True) -> np.ndarray:
    """Literal group semantics of hgnnaggr_cuda.cu:14-47, sequential (hg_oracle.c)."""
    key, row, st, ed, colind = map(_i32, (key, row, st, ed, colind))
    Xn = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
    Xn = np.ascontiguousarray(Xn, dtype=np.float32)
    N = int(num_nodes if num_nodes is not None else Xn.shape[0])
    F = Xn.shape[1]
    Y = np.zeros((N, F), np.float64 if f64 else np.float32)
    s1, s2, a_out, a_in = map(_f32, (s1, s2, a_out, a_in))
    lib().orc_aggr_groups(C.c_int64(row.size), C.c_int64(F), key.ctypes.data_as(C.c_void_p),
                          row.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p),
                          ed.ctypes.data_as(C.c_void_p), colind.ctypes.data_as(C.c_void_p),
                          Xn.ctypes.data_as(C.c_void_p), _opt(s1), _opt(s2), _opt(a_out), _opt(a_in),
                          Y.ctypes.data_as(C.c_void_p), C.c_int(1 if f64 else 0))
    return Y


def c_aggr_formula(t_indptr, t_indices, X, s1=None, s2=None, a_out=None, a_in=None,
                   num_nodes=None, reduce="sum") -> np.ndarray:
    """fp64 two-step formula of model/pygnn/hgnn.py:30-37 over the CSR of H_T (hg_oracle.c)."""
    t_indptr, t_indices = _i32(t_indptr), _i32(t_indices)
    Xn = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
    Xn = np.ascontiguousarray(Xn, dtype=np.float32)
    N = int(num_nodes if num_nodes is not None else Xn.shape[0])
    F = Xn.shape[1]
    Y = np.zeros((N, F), np.float64)
    s1, s2, a_out, a_in = map(_f32, (s1, s2, a_out, a_in))
    lib().orc_aggr_formula_f64(C.c_int64(N), C.c_int64(t_indptr.size - 1), C.c_int64(F),
                               t_indptr.ctypes.data_as(C.c_void_p),
                               t_indices.ctypes.data_as(C.c_void_p), Xn.ctypes.data_as(C.c_void_p),
                               _opt(s1), _opt(s2), _opt(a_out), _opt(a_in),
                               Y.ctypes.data_as(C.c_void_p), C.c_int({"sum": 0, "mean": 1}[reduce]))
    return Y


def c_hyperaggr_host(indptr, indices, t_indptr, t_indices, X) -> np.ndarray:
    """hg_oracle.c restatement of include/util/check.cuh:82-114 (un-scaled, fp32)."""
    indptr, indices, t_indptr, t_indices = map(_i32, (indptr, indices, t_indptr, t_indices))
    Xn = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
    N, F = indptr.size - 1, Xn.shape[1]
    Y = np.zeros((N, F), np.float32)
    lib().orc_hyperaggr_host(C.c_int64(N), C.c_int64(F), indptr.ctypes.data_as(C.c_void_p),
                             indices.ctypes.data_as(C.c_void_p), t_indptr.ctypes.data_as(C.c_void_p),
                             t_indices.ctypes.data_as(C.c_void_p), Xn.ctypes.data_as(C.c_void_p),
                             Y.ctypes.data_as(C.c_void_p))
    return Y


def ref_hyperaggr_host(indptr, indices, t_indptr, t_indices, X) -> np.ndarray:
    """The reference's own hyperaggr_reference_host (check.cuh:82-114), compiled in place."""
    indptr, indices, t_indptr, t_indices = map(_i32, (indptr, indices, t_indptr, t_indices))
    Xn = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
    N, F = indptr.size - 1, Xn.shape[1]
    Y = np.zeros((N, F), np.float32)
    ref_lib().ref_hyperaggr_host(C.c_int(N), C.c_int(F), indptr.ctypes.data_as(C.c_void_p),
                                 indices.ctypes.data_as(C.c_void_p),
                                 t_indptr.ctypes.data_as(C.c_void_p),
                                 t_indices.ctypes.data_as(C.c_void_p), Xn.ctypes.data_as(C.c_void_p),
                                 Y.ctypes.data_as(C.c_void_p))
    return Y


def ref_lab_gpu(variant: int, ngs: int, num_edges: int, key, st, ed, t_indices, X, out=None, iters: int = 0):
    """The reference's own lab kernels on the GPU, compiled in place for sm_100a (oracle/ref_shim.cu):
    variant 0 = HyperGAggr_Edgefused_Balance_Full_Kernel (`edge_based_full`, include/hgnnAgg.cuh:98-131,
    launch geometry :985-1003), 1 = ..._Shm_Kernel (`edge_based_shm`, :170-276, :1004-1017).  Un-scaled
    operator Y = H H^T X over the balancer groups; all arguments are CUDA tensors (int32 / float32).
    F must be < 32 or a multiple of 32 (the reference's own restriction).  With iters > 0 returns
    (Y, microseconds per call) timed the way the extension runs (zero-fill + kernel, hgnnaggr_cuda.cu:374),
    back to back on the legacy default stream."""
    lib_ = ref_lib()
    fn = lib_.ref_lab_full_gpu
    fn.restype = C.c_int
    fn.argtypes = [C.c_int] * 5 + [C.c_void_p] * 6
    F = int(X.shape[1])
    Y = torch.zeros_like(X) if out is None else out

    def launch():
        Y.zero_()
        rc = fn(int(variant), int(num_edges), int(ngs), int(st.numel()), F, key.data_ptr(), st.data_ptr(),
                ed.data_ptr(), t_indices.data_ptr(), X.data_ptr(), Y.data_ptr())
        if rc != 0:
            raise RuntimeError(f"reference lab kernel launch failed: cudaError {rc}")
    torch.cuda.synchronize()
    launch()
    torch.cuda.synchronize()
    if iters <= 0:
        return Y
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        launch()
    b.record()
    torch.cuda.synchronize()
    return Y, a.elapsed_time(b) / iters * 1e3


def ref_weight_grad(t_indptr, t_indices, G, X) -> np.ndarray:
    """The reference's hgnnbp_reference_host (check.cuh:116-143), compiled in place."""
    t_indptr, t_indices = _i32(t_indptr), _i32(t_indices)
    Gn = np.ascontiguousarray(np.asarray(G, dtype=np.float32))
    Xn = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
    M, F = t_indptr.size - 1, Xn.shape[1]
    out = np.zeros(M, np.float32)
    ref_lib().ref_weight_grad(C.c_int(M), C.c_int(F), t_indptr.ctypes.data_as(C.c_void_p),
                              t_indices.ctypes.data_as(C.c_void_p), Gn.ctypes.data_as(C.c_void_p),
                              Xn.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def c_weight_grad(t_indptr, t_indices, X, G, s1=None, a_out=None, a_in=None) -> np.ndarray:
    t_indptr, t_indices = _i32(t_indptr), _i32(t_indices)
    Gn = np.ascontiguousarray(np.asarray(G, dtype=np.float32))
    Xn = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
    M, F = t_indptr.size - 1, Xn.shape[1]
    out = np.zeros(M, np.float64)
    s1, a_out, a_in = map(_f32, (s1, a_out, a_in))
    lib().orc_weight_grad_f64(C.c_int64(M), C.c_int64(F), t_indptr.ctypes.data_as(C.c_void_p),
                              t_indices.ctypes.data_as(C.c_void_p), Xn.ctypes.data_as(C.c_void_p),
                              Gn.ctypes.data_as(C.c_void_p), _opt(s1), _opt(a_out), _opt(a_in),
                              out.ctypes.data_as(C.c_void_p))
    return out


def c_aggr_max(t_indptr, t_indices, X, s1=None, s2=None, a_out=None, num_nodes=None):
    """f1-max forward (hgnnaggr_cuda.cu:144-178): returns (Y fp64 [N,F], record int32 [M,F])."""
    t_indptr, t_indices = _i32(t_indptr), _i32(t_indices)
    Xn = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
    N = int(num_nodes if num_nodes is not None else Xn.shape[0])
    M, F = t_indptr.size - 1, Xn.shape[1]
    Y = np.zeros((N, F), np.float64)
    rec = np.zeros((M, F), np.int32)
    s1, s2, a_out = map(_f32, (s1, s2, a_out))
    lib().orc_aggr_max_fwd(C.c_int64(N), C.c_int64(M), C.c_int64(F),
                           t_indptr.ctypes.data_as(C.c_void_p), t_indices.ctypes.data_as(C.c_void_p),
                           Xn.ctypes.data_as(C.c_void_p), _opt(s1), _opt(s2), _opt(a_out),
                           Y.ctypes.data_as(C.c_void_p), rec.ctypes.data_as(C.c_void_p))
    return Y, rec


def c_aggr_max_bwd(t_indptr, t_indices, G, record, s1=None, s2=None, a_out=None, num_nodes=None):
    """f1-max backward (hgnnaggr_cuda.cu:180-208)."""
    t_indptr, t_indices = _i32(t_indptr), _i32(t_indices)
    Gn = np.ascontiguousarray(np.asarray(G, dtype=np.float32))
    N = int(num_nodes if num_nodes is not None else Gn.shape[0])
    M, F = t_indptr.size - 1, Gn.shape[1]
    rec = np.ascontiguousarray(np.asarray(record, dtype=np.int32))
    dX = np.zeros((N, F), np.float64)
    s1, s2, a_out = map(_f32, (s1, s2, a_out))
    lib().orc_aggr_max_bwd(C.c_int64(N), C.c_int64(M), C.c_int64(F),
                           t_indptr.ctypes.data_as(C.c_void_p), t_indices.ctypes.data_as(C.c_void_p),
                           Gn.ctypes.data_as(C.c_void_p), _opt(s1), _opt(s2), _opt(a_out),
                           rec.ctypes.data_as(C.c_void_p), dX.ctypes.data_as(C.c_void_p))
    return dX


def torch_hgnn_conv(X, V, E, degE, degV, W, num_nodes, num_edges, first_aggr="sum"):
    """Pure-torch restatement of the PyG back-end conv, model/pygnn/hgnn.py:30-37
    (torch_scatter is not installed, so ``scatter(..., reduce)`` is ``index_add_``).
    This is the 'PyG-equivalent CPU path' bench.py times; works in any float dtype."""
    Xve = X[V]                                                            # :30
    Xe = torch.zeros(num_edges, X.shape[1], dtype=X.dtype, device=X.device).index_add_(0, E, Xve)
    if first_aggr == "mean":                                              # :31 reduce=first_aggr
        cnt = torch.zeros(num_edges, dtype=X.dtype, device=X.device).index_add_(
            0, E, torch.ones_like(E, dtype=X.dtype))
        Xe = Xe / cnt.clamp_(min=1).unsqueeze(1)
    if degE is not None:
        Xe = Xe * degE.reshape(-1, 1).to(X.dtype)                         # :32
    if W is not None:
        Xe = Xe * W.reshape(-1, 1).to(X.dtype)                            # :33
    Xev = Xe[E]                                                           # :34
    Xv = torch.zeros(num_nodes, X.shape[1], dtype=X.dtype, device=X.device).index_add_(0, V, Xev)
    if degV is not None:
        Xv = Xv * degV.reshape(-1, 1).to(X.dtype)                         # :36
    return Xv


def rel_err(got, want64) -> float:
    """max |got - want| / max(|want|) -- the 1e-5 'relative' figure of the north star,
    normalised by the largest magnitude so exact zeros do not blow it up."""
    got = np.asarray(got, dtype=np.float64)
    want64 = np.asarray(want64, dtype=np.float64)
    scale = max(float(np.abs(want64).max()) if want64.size else 0.0, 1e-30)
    return float(np.abs(got - want64).max() / scale) if want64.size else 0.0


def rel_err_terms(got, want64, bound64, eps: float = 1e-30) -> float:
    """Element-wise error scaled by the sum of the magnitudes of the terms of each output element
    (SURVEY.md 7.3-4): max_ij |got - want| / bound, bound_ij = sum |terms| (the same aggregation applied
    to |X| with |scales|).  Tighter than rel_err: an element is not excused by a large value elsewhere."""
    got = np.asarray(got, dtype=np.float64)
    want64 = np.asarray(want64, dtype=np.float64)
    bound64 = np.asarray(bound64, dtype=np.float64)
    if want64.size == 0:
        return 0.0
    return float((np.abs(got - want64) / np.maximum(bound64, eps)).max())


def ref_read_mtx(path):
    """The reference's own MatrixMarket loader (include/dataloader/dataloader.hpp:22-104), compiled in place:
    ``(nrow, ncol, indptr, indices, rowind)``.  Only for well-formed files -- it calls exit() otherwise."""
    lib_ = ref_lib()
    nrow, ncol, nnz = C.c_int(), C.c_int(), C.c_int()
    lib_.ref_read_mtx(os.fsencode(path), C.byref(nrow), C.byref(ncol), C.byref(nnz))
    indptr = np.empty(nrow.value + 1, np.int32)
    indices = np.empty(nnz.value, np.int32)
    rowind = np.empty(nnz.value, np.int32)
    lib_.ref_read_mtx_fetch(indptr.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p),
                            rowind.ctypes.data_as(C.c_void_p))
    return nrow.value, ncol.value, indptr, indices, rowind
