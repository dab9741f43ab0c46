"""Locality reordering: the permutation is a relabelling -- the operator commutes with it -- and
reverse Cuthill-McKee brings back the locality a shuffled numbering destroyed.  Host logic; runs without a GPU."""
import numpy as np
import torch

import hypergef_b200 as hgef
from hypergef_b200 import reorder, synth
from oracle import oracle as orc


def _shuffled(data, seed):
    N, M = data.num_nodes, data.num_hyperedges
    g = torch.Generator().manual_seed(seed)
    return reorder.permute_data(data, torch.randperm(N, generator=g), torch.randperm(M, generator=g))


def test_operator_commutes_with_the_permutation():
    data = synth.make_shape("cora", seed=3, num_feat=8)
    N = data.num_nodes
    new, vperm, eperm = reorder.reorder(_shuffled(data, 1))
    for d in (data, new):
        assert d.edge_index.shape == data.edge_index.shape
    assert sorted(vperm.tolist()) == list(range(N)) and sorted(eperm.tolist()) == list(range(data.num_hyperedges))

    def aggregate(d):
        hg = hgef.HyperGraph(d, torch.device("cpu"), "cora")
        return orc.c_aggr_formula(hg.H_T_csrptr.numpy(), hg.H_T_colind.numpy(), d.x.numpy(),
                                  s1=hg.degE.numpy(), a_out=hg.degV.numpy())
    shuf = _shuffled(data, 1)
    want = aggregate(shuf)                                   # in the shuffled numbering
    got = aggregate(new)                                     # in the reordered numbering
    assert orc.rel_err(reorder.restore_rows(torch.from_numpy(got), vperm).numpy(), want) < 1e-6
    # features and labels moved with their vertices
    assert torch.equal(reorder.restore_rows(new.x, vperm), shuf.x) and torch.equal(reorder.restore_rows(new.y, vperm), shuf.y)


def test_rcm_recovers_block_locality():
    data = synth.make_shape("pubmed", replicas=8, seed=0)       # 8 disjoint replicas: spans stay inside a replica
    N = data.num_nodes
    natural = reorder.mean_span(data.edge_index, N)
    shuf = _shuffled(data, 2)
    shuffled = reorder.mean_span(shuf.edge_index, N)
    new, _, _ = reorder.reorder(shuf)
    restored = reorder.mean_span(new.edge_index, N)
    assert shuffled > 3 * natural                                # the shuffle spreads every hyperedge over the whole id range
    assert restored < 0.75 * natural, (natural, shuffled, restored)  # RCM does better than the generator's own numbering
