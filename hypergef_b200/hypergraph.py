"""``HyperGraph`` -- drop-in for ``HyperGsys/hypergraph.py:10-101`` (minus the dgl parts).

Same constructor ``HyperGraph(data, device, data_name)`` and the same attributes the
aggregation path reads (``H_csrptr, H_colind, H_data, H_T_csrptr, H_T_colind, H_T_data,
V, E, degV, degE, degD, group_key, group_row, group_start, group_end, num_nodes,
num_edges, nnz, adj_g1, adj_g2``), with contents bit-identical to the scipy route of the
reference, but built natively:

* CSR of ``H`` and ``H^T``  -> ``hg_csr_build_dev`` (radix sort + run-length encode on the GPU)
  or ``hg_csr_build_host`` when ``device`` is the CPU; scipy semantics (sorted columns,
  duplicates summed) -- ``hypergraph.py:23-25``;
* ``degV = rowsum^-1/2`` with ``inf -> 1``, ``degE = colsum^-1`` -- ``hypergraph.py:34-45``;
* the balancer arrays -> :class:`hypergef_b200.balancer.balance_schedule`, stored as int32
  directly (the reference goes through ``torch.Tensor(list)``, a float32 round trip that
  corrupts offsets >= 2**24, ``hypergraph.py:98-101``).

``ngs`` comes from the reference's per-dataset table (``hypergraph.py:74-75``); synthetic
graphs pass ``ngs=`` explicitly (or carry ``data.ngs``).
Not rebuilt: the dgl Laplacian ``self.L`` and ``dgl_prepare`` (comparison back-end only).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native
from .balancer import balance_schedule

__all__ = ["HyperGraph", "PARTITION_DICT"]

# hypergraph.py:74-75
PARTITION_DICT = {"yelp": 400, "20newsW100": 400, "coauthor_cora": 10, "zoo": 20, "NTU2012": 80,
                  "cora": 210, "pubmed": 40, "Mushroom": 250, "coauthor_dblp": 80,
                  "house-committees": 40, "walmart-trips": 210, "citeseer": 6, "ModelNet40": 300}


def _dev_index(device: torch.device) -> int:
    return device.index if device.index is not None else torch.cuda.current_device()


def build_csr(V: torch.Tensor, E: torch.Tensor, num_nodes: int, num_edges: int, device: torch.device):
    """(H_csrptr, H_colind, H_data, H_T_csrptr, H_T_colind, H_T_data) on ``device``."""
    nnz = V.numel()
    V = V.to(device=device, dtype=torch.int64).contiguous()
    E = E.to(device=device, dtype=torch.int64).contiguous()
    i32 = dict(dtype=torch.int32, device=device)
    indptr, t_indptr = torch.empty(num_nodes + 1, **i32), torch.empty(num_edges + 1, **i32)
    indices, t_indices = torch.empty(nnz, **i32), torch.empty(nnz, **i32)
    data, t_data = (torch.empty(nnz, dtype=torch.float32, device=device) for _ in range(2))
    out = C.c_int64()
    args = (num_nodes, num_edges, nnz, V.data_ptr(), E.data_ptr(), indptr.data_ptr(), indices.data_ptr(),
            data.data_ptr(), t_indptr.data_ptr(), t_indices.data_ptr(), t_data.data_ptr(), C.byref(out))
    if device.type == "cuda":
        dev = _dev_index(device)
        _native.call("hg_csr_build_dev", *args, dev, torch.cuda.current_stream(dev).cuda_stream)
    else:
        _native.call("hg_csr_build_host", *args)
    z = out.value
    if z != nnz:  # duplicate pairs were merged (scipy sum_duplicates)
        indices, data, t_indices, t_data = (a[:z].clone() for a in (indices, data, t_indices, t_data))
    return indptr, indices, data, t_indptr, t_indices, t_data


def degree_scale(indptr: torch.Tensor, data: torch.Tensor, power: float, inf_to_one: bool) -> torch.Tensor:
    """``(row sums)^power`` as an ``[n, 1]`` float32 tensor (hypergraph.py:34-45)."""
    n = indptr.numel() - 1
    if indptr.is_cuda:
        out = torch.empty((n, 1), dtype=torch.float32, device=indptr.device)
        dev = _dev_index(indptr.device)
        _native.call("hg_degree_scale_dev", n, indptr.data_ptr(), data.data_ptr(), float(power),
                     int(inf_to_one), out.data_ptr(), dev, torch.cuda.current_stream(dev).cuda_stream)
        return out
    # host graph construction (the reference does all of it on the host)
    rows = torch.repeat_interleave(torch.arange(n), (indptr[1:] - indptr[:-1]).long())
    deg = torch.zeros(n, dtype=torch.float64).index_add_(0, rows, data.double()).float().unsqueeze(1)
    out = deg.pow(power)
    if inf_to_one:
        out[torch.isinf(out)] = 1
    return out


class HyperGraph:
    def __init__(self, data, device, data_name, ngs=None):
        self.device = torch.device(device)
        self.data_name = data_name
        self.num_nodes = int(data.x.shape[0])
        edge_index = data.edge_index
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("data.edge_index must be [2, 2*nnz]")
        hit = torch.where(edge_index[0] == self.num_nodes)[0]
        if hit.numel() == 0:
            raise ValueError("edge_index has no hyperedge->vertex half (no row id == num_nodes); "
                             "expected [[V ; E+N], [E+N ; V]] (hypergraph.py:15)")
        c_idx = int(hit.min())
        V2E = edge_index[:, :c_idx]
        V = V2E[0]
        E = V2E[1] - self.num_nodes
        self.num_edges = int(torch.unique(V2E[1]).numel())          # hypergraph.py:19
        self.nnz = int(V2E.shape[1])                                # hypergraph.py:20 (before merging)
        if self.nnz and (int(E.min()) < 0 or int(E.max()) >= self.num_edges):
            raise ValueError("hyperedge ids must be consecutive 0..num_edges-1 after subtracting "
                             "num_nodes (scipy raises on the same input, hypergraph.py:24)")

        (self.H_csrptr, self.H_colind, self.H_data, self.H_T_csrptr, self.H_T_colind,
         self.H_T_data) = build_csr(V, E, self.num_nodes, self.num_edges, self.device)
        self.V, self.E = V.to(self.device), E.to(self.device)

        degV = degree_scale(self.H_csrptr, self.H_data, -0.5, False)           # [N,1]
        self.degE = degree_scale(self.H_T_csrptr, self.H_T_data, -1.0, False)  # [M,1]
        self.degD = degV.pow(-1)            # hypergraph.py:42: before the inf fix (isolated -> 0)
        degV[torch.isinf(degV)] = 1         # hypergraph.py:45
        self.degV = degV

        self.adj_g1 = self.H_csrptr, self.H_colind, self.H_data
        self.adj_g2 = self.H_T_csrptr, self.H_T_colind, self.H_T_data

        if ngs is None:
            ngs = getattr(data, "ngs", None)
        if ngs is None:
            if data_name not in PARTITION_DICT:
                raise KeyError(f"no partition size for dataset {data_name!r}; pass ngs= "
                               f"(known: {sorted(PARTITION_DICT)})")
            ngs = PARTITION_DICT[data_name]
        self.ngs = int(ngs)
        self.balance(self.ngs, self.H_T_csrptr)

    def balance(self, ngs, H_T_csrptr):
        """hypergraph.py:96-101, int32 all the way."""
        bs = balance_schedule(ngs, H_T_csrptr)
        as_dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(a)).to(self.device)
        self.group_start = as_dev(bs.group_st)
        self.group_end = as_dev(bs.group_ed)
        self.group_key = as_dev(bs.balan_key)
        self.group_row = as_dev(bs.balan_row)

    # scipy views for callers that want them (store_mtx in the reference, hypergraph.py:79-85)
    def _scipy(self, indptr, indices, data, shape):
        import scipy.sparse as sp
        return sp.csr_matrix((data.cpu().numpy().astype("float64"), indices.cpu().numpy(),
                              indptr.cpu().numpy()), shape=shape)

    @property
    def H(self):
        return self._scipy(self.H_csrptr, self.H_colind, self.H_data, (self.num_nodes, self.num_edges))

    @property
    def H_T(self):
        return self._scipy(self.H_T_csrptr, self.H_T_colind, self.H_T_data, (self.num_edges, self.num_nodes))

    def store_mtx(self, path):
        from scipy.io import mmwrite
        mmwrite(path + self.data_name + ".mtx", self.H)
