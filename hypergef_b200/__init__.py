"""hypergef_b200 -- B200-native fused hypergraph aggregation (HyperGef's hot path).

Public surface (names follow the reference):

    HyperGraph, balance_schedule                     graph + balancer construction
    HGNNAggr, UniGNNConvdeg, UniGNNConv              Python op wrappers
    hgnnaggr, unignnaggr                             stand-ins for the two torch extensions
    convs                                            hgsys conv layers / 2-layer HGNN
    io.read_mtx, io.hypergraph_from_mtx              MatrixMarket incidence files
    compat.install()                                 register the reference's module names

Everything computes through ``libhgef_b200.so`` (C-ABI, include/hgef_b200.h); nothing here
falls back to the CPU or to eager PyTorch.
"""
from . import _native, hgnnaggr, io, ops, unignnaggr  # noqa: F401
from .balancer import balance_schedule  # noqa: F401
from .hypergraph import PARTITION_DICT, HyperGraph  # noqa: F401
from .ops import get_backward_mode, set_backward_mode  # noqa: F401
from .wrappers import HGNNAggr, UniGNNConv, UniGNNConvdeg  # noqa: F401

__version__ = "0.1.0"
