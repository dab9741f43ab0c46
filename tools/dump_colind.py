#!/usr/bin/env python
"""Write the H^T column-index list of a synthetic shape as raw int32 (input of tools/replay.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hypergef_b200 as hgef
from hypergef_b200 import synth
shape, reps, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
dev = torch.device("cuda:0")
data = synth.make_shape(shape, replicas=reps, seed=0, device=dev)
hg = hgef.HyperGraph(data, dev, data.dataset)
hg.H_T_colind.cpu().numpy().astype("int32").tofile(out)
print("wrote", out, hg.H_T_colind.numel())
