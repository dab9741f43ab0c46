"""hgsys conv layers and the 2-layer model the epoch-ms metric is quoted on.

Signatures follow ``HyperGsys/model/ugsys/{hgnn,unigin,unigcnii}.py`` and
``HyperGsys/model/gnn.py:110-134``; only the aggregation inside is new.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .wrappers import HGNNAggr, UniGNNConv, UniGNNConvdeg

__all__ = ["HyperGsysHGNN", "HyperGsysUinGINConv", "HyperGsysUniGCNIIConv", "HGsysHGNN", "ColumnParallelHGNN"]


class HyperGsysHGNN(nn.Module):
    """model/ugsys/hgnn.py:7-27: ``Linear`` then the fused HGNN aggregation."""


    def __init__(self, hyperg, in_channels, out_channels, first_aggr="sum", heads=1, project="vertex"):
        super().__init__()
        # project: "vertex" = the reference's order (Linear on the N vertex rows, then the aggregation);
        # "edge" / "after" / "auto" = SURVEY.md 8(f) N1, see ops.projected_aggregate
        self.project = project
        self.W = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.Wdiag = torch.ones(hyperg.degE.shape[0], device=hyperg.device)
        self.heads, self.in_channels, self.out_channels = heads, in_channels, out_channels
        self.hyperg, self.degE, self.degV = hyperg, hyperg.degE, hyperg.degV
        self.first_aggr = first_aggr

    def forward(self, X):
        if self.project != "vertex" and self.first_aggr in ("sum", None):
            # the same layer with the projection moved to where it is cheapest (ops.projection_order)
            return ops.projected_aggregate(self.hyperg, X, self.W.weight.t(), self.degE, self.degV, self.Wdiag,
                                           order=self.project)
        X = self.W(X)
        return HGNNAggr(self.hyperg, X, self.degE, self.degV, self.Wdiag, self.first_aggr)


class HyperGsysUinGINConv(nn.Module):
    """model/ugsys/unigin.py:7-26: ``(1+eps) XW + H H^T XW`` (reference spelling kept)."""

    def __init__(self, hyperg, in_channels, out_channels, first_aggr="sum", heads=1):
        super().__init__()
        self.W = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.heads, self.in_channels, self.out_channels = heads, in_channels, out_channels
        self.hyperg, self.degE, self.degV = hyperg, hyperg.degE, hyperg.degV
        self.eps = nn.Parameter(torch.zeros(1))

    def forward(self, X):
        X = self.W(X)
        return (1 + self.eps) * X + UniGNNConv(self.hyperg, X)


class HyperGsysUniGCNIIConv(nn.Module):
    """model/ugsys/unigcnii.py:7-26 with the ``alpha`` / ``beta`` it reads but never defines
     passed at call time, as the pyg twin does (model/pygnn/unigcnii.py)."""

    def __init__(self, hyperg, in_channels, out_channels, first_aggr="sum", heads=1):
        super().__init__()
        self.W = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.hyperg, self.degE, self.degV = hyperg, hyperg.degE, hyperg.degV

    def forward(self, X, X0, alpha, beta):
        Xv = UniGNNConvdeg(self.hyperg, X, self.degE, self.degV)
        Xi = (1 - alpha) * Xv + alpha * X0
        return (1 - beta) * Xi + beta * self.W(Xi)


class HGsysHGNN(nn.Module):
    """model/gnn.py:110-134: ``nlayer-1`` hidden convs + ``conv_out``, ReLU, dropouts, log-softmax."""

    def __init__(self, args, hyperg, nfeat, nhid, nclass, nlayer=2, first_aggr="sum", nhead=1,
                 conv=HyperGsysHGNN):
        super().__init__()
        args = args or SimpleNamespace(activation="relu", input_drop=0.6, dropout=0.6)
        self.conv_out = conv(hyperg, nhid * nhead, nclass, first_aggr, nhead)
        self.convs = nn.ModuleList(
            [conv(hyperg, nfeat, nhid, first_aggr, nhead)] +
            [conv(hyperg, nhid * nhead, nhid, first_aggr, nhead) for _ in range(nlayer - 2)])
        self.act = {"relu": nn.ReLU(), "leaky_relu": nn.LeakyReLU()}[args.activation]
        self.input_drop = nn.Dropout(args.input_drop)
        self.dropout = nn.Dropout(args.dropout)

    def forward(self, X):
        X = self.input_drop(X)
        for conv in self.convs:
            X = self.dropout(self.act(conv(X)))
        return F.log_softmax(self.conv_out(X), dim=1)


class _GatherColumns(torch.autograd.Function):
    """all_gather of column blocks ``[N, F/P] -> [N, F]`` whose consumers are REPLICATED on every rank (each rank
    then computes the same loss): the gradient of the local block is the matching column slice of the incoming
    gradient, with no reduction across ranks."""

    @staticmethod
    def forward(ctx, x, group):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        ctx.rank, ctx.cols = dist.get_rank(group), x.shape[1]
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, dim=1)

    @staticmethod
    def backward(ctx, g):
        return g[:, ctx.rank * ctx.cols:(ctx.rank + 1) * ctx.cols].contiguous(), None


class ColumnParallelHGNN(nn.Module):
    """The 2-layer HGNN of ``model/gnn.py:110-134`` over ``P`` GPUs by FEATURE-COLUMN sharding:
    the aggregation never mixes columns (``hgnnaggr_cuda.cu:21,34,44``), so rank r holds the replicated graph,
    the replicated input features and the column block ``W1[:, r]`` of the first layer, aggregates its own
    ``nhid / P`` hidden columns with no collective, and one ``all_gather`` of ``[N, nhid / P]`` activations
    feeds the (small, replicated) output layer.  The reference is single-GPU (``hgsys.py:54``)."""

    def __init__(self, hyperg, nfeat, nhid, nclass, group=None, input_drop=0.6, dropout=0.6):
        super().__init__()
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if nhid % self.world:
            raise ValueError(f"nhid={nhid} is not divisible by the {self.world} ranks")
        self.conv1 = HyperGsysHGNN(hyperg, nfeat, nhid // self.world)
        self.conv_out = HyperGsysHGNN(hyperg, nhid, nclass)
        self.input_drop, self.dropout = nn.Dropout(input_drop), nn.Dropout(dropout)

    def forward(self, X):
        h = torch.relu(self.conv1(self.input_drop(X)))
        if self.world > 1:
            h = _GatherColumns.apply(h, self.group)
        return F.log_softmax(self.conv_out(self.dropout(h)), dim=1)
