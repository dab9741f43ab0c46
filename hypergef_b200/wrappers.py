"""Python op wrappers -- drop-in for ``HyperGsys/source/python/{hgnnaggr,unignnconv}.py``.

``hyperg`` / ``dl`` is duck-typed on ``group_key, group_row, group_start, group_end,
H_T_csrptr, H_T_colind`` exactly as in the reference.
"""
from __future__ import annotations

from . import ops

__all__ = ["HGNNAggr", "UniGNNConvdeg", "UniGNNConv"]


def HGNNAggr(hyperg, in_feat, degE, degV, Wdiag, first_aggr="sum"):
    """source/python/hgnnaggr.py:6-7.  The reference drops ``first_aggr``; here ``'mean'`` and
    ``'max'`` reach the first-stage variants the extension already exports (hgnnaggr.cc:149-150),
    and the reference test's 5-argument call (test/hgnn_test.py:89) works through the default."""
    if first_aggr in ("sum", None):
        return ops.hgnnaggr(hyperg.group_key, hyperg.group_row, hyperg.group_start, hyperg.group_end,
                            hyperg.H_T_csrptr, hyperg.H_T_colind, in_feat, degE, degV, Wdiag)
    if first_aggr == "mean":
        return ops.hgnnaggr_mean(hyperg.H_T_csrptr, hyperg.H_T_colind, in_feat, degE, degV, Wdiag)
    if first_aggr == "max":
        return ops.hgnnaggr_max(hyperg.H_T_csrptr, hyperg.H_T_colind, in_feat, degE, degV, Wdiag)[0]
    raise ValueError(f"first_aggr must be 'sum', 'mean' or 'max', got {first_aggr!r}")


def UniGNNConvdeg(dl, in_feat, degE, degV):
    """source/python/unignnconv.py:6-7."""
    return ops.unignnaggrdeg(dl.group_key, dl.group_row, dl.group_start, dl.group_end, dl.H_T_csrptr,
                             dl.H_T_colind, in_feat, degE, degV)


def UniGNNConv(dl, in_feat):
    """source/python/unignnconv.py:9-10."""
    return ops.unignnaggr(dl.group_key, dl.group_row, dl.group_start, dl.group_end, dl.H_T_csrptr,
                          dl.H_T_colind, in_feat)
