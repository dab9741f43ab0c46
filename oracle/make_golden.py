"""Generate tests/golden/*.npz by RUNNING THE REFERENCE in the build container.

Run once here (``python oracle/make_golden.py``); the fixtures are committed
because /root/reference does not exist on the GPU box.  Sources of truth:

* balancer vectors      -- the reference's own ``HyperGsys/balancer.py`` imported by
                           file path (needs only torch + numpy) and converted exactly as
                           ``HyperGsys/hypergraph.py:96-101`` does (``torch.Tensor(list).int()``);
                           cross-checked against the C++ twin compiled into oracle/_ref.
* CSR / degree vectors  -- scipy, called exactly as ``HyperGsys/hypergraph.py:15-49`` calls it
                           (the file itself cannot be imported: it needs dgl at import time).
* aggregation vectors   -- the reference's host golden ``hyperaggr_reference_host``
                           (``include/util/check.cuh:82-114``) compiled in place (oracle/_ref),
                           plus the fp64 two-step formula of ``model/pygnn/hgnn.py:30-37``.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("HG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import oracle as orc                     # noqa: E402
from hypergef_b200 import synth                      # noqa: E402


def load_reference_balancer():
    spec = importlib.util.spec_from_file_location("ref_balancer", os.path.join(REF, "HyperGsys", "balancer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.balance_schedule


def run_reference_balancer(balance_schedule, ngs, csrptr_np):
    bs = balance_schedule(ngs, torch.from_numpy(np.asarray(csrptr_np, dtype=np.int32)))
    # hypergraph.py:98-101
    conv = lambda lst: torch.Tensor(lst).int().numpy()
    return dict(key=conv(bs.balan_key), row=conv(bs.balan_row), st=conv(bs.group_st), ed=conv(bs.group_ed))


def main():
    os.makedirs(OUT, exist_ok=True)
    orc.build(ref=True)
    balance_schedule = load_reference_balancer()
    rng = np.random.default_rng(0)

    # ---- balancer: hand cases + random csrptrs ----
    cases = {
        "toy_a": (3, [0, 5, 5, 7, 14]),
        "toy_b": (3, [0, 4, 4]),
        "toy_c": (3, [0, 6]),
        "exact_multiple": (4, [0, 4, 8, 16, 16, 20]),
        "ngs1": (1, [0, 2, 3, 3, 6]),
        "ngs_big": (1000, [0, 1, 3, 6, 10]),
        "leading_empty": (2, [0, 0, 0, 5, 6]),
        "trailing_empty": (2, [0, 5, 6, 6, 6]),
    }
    for i in range(6):
        nrow = int(rng.integers(1, 60))
        deg = rng.integers(0, 40, size=nrow)
        if i % 2:
            deg[rng.integers(0, nrow)] = int(rng.integers(100, 700))
        if deg.sum() == 0:
            deg[0] = 1
        cases[f"rand{i}"] = (int(rng.integers(1, 17)), np.concatenate([[0], np.cumsum(deg)]).tolist())
    bal = {}
    for name, (ngs, ptr) in cases.items():
        got = run_reference_balancer(balance_schedule, ngs, ptr)
        twin = orc.ref_balancer(ngs, ptr)
        for k, a in (("key", twin.balan_key), ("row", twin.balan_row), ("st", twin.group_st), ("ed", twin.group_ed)):
            assert np.array_equal(got[k], a), f"reference .py and C++ twin disagree on {name}:{k}"
        bal[f"{name}__ngs"] = np.int32(ngs)
        bal[f"{name}__csrptr"] = np.asarray(ptr, np.int32)
        for k, a in got.items():
            bal[f"{name}__{k}"] = a
    np.savez_compressed(os.path.join(OUT, "balancer.npz"), **bal)
    print("balancer.npz:", len(cases), "cases")

    # ---- graph construction + degrees + balancer on synthetic shapes ----
    small = synth.HyperShape("mini", 300, 120, 6, "zipf", min_size=1, max_size=90, zipf_alpha=1.7, force_max=75)
    graphs = {
        "cora": (synth.SHAPES["cora"], 1, 210),
        "mini": (small, 1, 6),
        "mini_rep3": (small, 3, 4),
    }
    for gname, (shape, reps, ngs) in graphs.items():
        V, E, N, M = synth.draw_incidence(shape, replicas=reps, seed=7)
        ei = synth.incidence_to_edge_index(V, E, N)
        if gname == "mini":
            # a duplicated incidence pair and an isolated vertex: scipy sums the duplicate
            # (data 2.0, degree counts it twice) and hypergraph.py:45 maps inf -> 1
            keep = V != 5
            V, E = V[keep], E[keep]
            E = torch.unique(E, return_inverse=True)[1]       # keep hyperedge ids consecutive
            M = int(E.max()) + 1
            dup = torch.stack((V[:1], E[:1] + N))
            first = torch.cat((dup, torch.stack((V, E + N))), dim=1)
            second = first.flip(0)
            second = second[:, torch.argsort(second[0] * (N + M) + second[1], stable=True)]
            ei = torch.cat((first, second), dim=1)
        Vr, Er, num_edges, nnz = orc.split_edge_index(ei, N)
        H, H_T = orc.scipy_incidence(Vr.numpy(), Er.numpy(), N, num_edges)
        degV, degE = orc.scipy_degrees(H)
        b = run_reference_balancer(balance_schedule, ngs, H_T.indptr)
        g = dict(edge_index=ei.numpy(), num_nodes=np.int64(N), num_edges=np.int64(num_edges), nnz=np.int64(nnz),
                 ngs=np.int32(ngs), H_csrptr=H.indptr.astype(np.int32), H_colind=H.indices.astype(np.int32),
                 H_data=H.data.astype(np.float32), H_T_csrptr=H_T.indptr.astype(np.int32),
                 H_T_colind=H_T.indices.astype(np.int32), H_T_data=H_T.data.astype(np.float32),
                 degV=degV.numpy(), degE=degE.numpy(), group_key=b["key"], group_row=b["row"],
                 group_start=b["st"], group_end=b["ed"])
        if gname != "cora":
            # aggregation goldens (small F so the fixture stays small)
            F = 8
            X = torch.randn(N, F, generator=torch.Generator().manual_seed(3)).numpy()
            g["X"] = X
            g["Y_unscaled_ref_host"] = orc.ref_hyperaggr_host(H.indptr, H.indices, H_T.indptr, H_T.indices, X)
            Wd = (0.5 + torch.rand(num_edges, generator=torch.Generator().manual_seed(4))).numpy()
            g["W"] = Wd.astype(np.float32)
            g["Y_hgnn_f64"] = orc.c_aggr_formula(H_T.indptr, H_T.indices, X, degE.numpy(), Wd, degV.numpy())
            # the pure-torch PyG-equivalent conv in fp64 must agree with the C formula
            Vt, Et = torch.from_numpy(H.tocoo().row.astype(np.int64)), torch.from_numpy(H.tocoo().col.astype(np.int64))
            Yt = orc.torch_hgnn_conv(torch.from_numpy(X).double(), Vt, Et, degE.double(), degV.double(),
                                     torch.from_numpy(Wd).double(), N, num_edges)
            if gname == "mini":
                # duplicates: the pattern-only kernels count the pair once, PyG's V/E lists twice
                pass
            else:
                assert np.allclose(Yt.numpy(), g["Y_hgnn_f64"], rtol=1e-12, atol=1e-12)
        np.savez_compressed(os.path.join(OUT, f"graph_{gname}.npz"), **g)
        print(f"graph_{gname}.npz: N={N} M={num_edges} nnz={nnz} S={b['key'].size - 1} G={b['row'].size}")


if __name__ == "__main__":
    main()
