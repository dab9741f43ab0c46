"""Register this package under the reference's module names.

After ``hypergef_b200.compat.install()`` the reference's own import lines resolve to the
B200 path without editing them:

    import hgnnaggr, unignnaggr                                   (source/python/*.py:3)
    from HyperGsys.balancer import balance_schedule               (hypergraph.py:7)
    from HyperGsys.hypergraph import HyperGraph                   (dataloader.py)
    from HyperGsys.source.python.hgnnaggr import HGNNAggr         (model/ugsys/hgnn.py:1)
    from HyperGsys.source.python.unignnconv import UniGNNConv     (model/ugsys/unigin.py:1)
"""
from __future__ import annotations

import sys
import types

from . import balancer, hgnnaggr, hypergraph, unignnaggr, wrappers


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    return mod


def install(force: bool = False) -> None:
    table = {
        "hgnnaggr": hgnnaggr,
        "unignnaggr": unignnaggr,
        "HyperGsys": _module("HyperGsys", __path__=[]),
        "HyperGsys.balancer": balancer,
        "HyperGsys.hypergraph": hypergraph,
        "HyperGsys.source": _module("HyperGsys.source", __path__=[]),
        "HyperGsys.source.python": _module("HyperGsys.source.python", __path__=[]),
        "HyperGsys.source.python.hgnnaggr": _module("HyperGsys.source.python.hgnnaggr",
                                                    HGNNAggr=wrappers.HGNNAggr, hgnnaggr=hgnnaggr),
        "HyperGsys.source.python.unignnconv": _module("HyperGsys.source.python.unignnconv",
                                                      UniGNNConvdeg=wrappers.UniGNNConvdeg,
                                                      UniGNNConv=wrappers.UniGNNConv, unignnaggr=unignnaggr),
    }
    for name, mod in table.items():
        if force or name not in sys.modules:
            sys.modules[name] = mod
    hs = sys.modules["HyperGsys"]
    if isinstance(hs, types.ModuleType) and not hasattr(hs, "balancer"):
        hs.balancer, hs.hypergraph = balancer, hypergraph
