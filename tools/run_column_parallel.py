#!/usr/bin/env python
"""ColumnParallelHGNN (feature-column sharded 2-layer HGNN) against the single-GPU HGsysHGNN with the same weights:
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/run_column_parallel.py
Every rank builds the full model from the same seed, takes its column block of the first layer, and compares output,
loss and gradients (dropout off).  Prints one JSON line on rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as Fn
import hypergef_b200 as hgef
from hypergef_b200 import convs, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
data = synth.make_shape("cora", seed=0)
hg = hgef.HyperGraph(data, dev, "cora")
N, nfeat, nhid, ncls = hg.num_nodes, 64, 32, 7
torch.manual_seed(4)                                      # the same full model and inputs on every rank
full = convs.HGsysHGNN(None, hg, nfeat, nhid, ncls).to(dev).eval()
X = torch.randn(N, nfeat, device=dev)
y = torch.randint(0, ncls, (N,), device=dev)
want = full(X)
wloss = Fn.nll_loss(want, y)
wloss.backward()
W1, W2 = full.convs[0].W.weight, full.conv_out.W.weight

par = convs.ColumnParallelHGNN(hg, nfeat, nhid, ncls).to(dev).eval()
blk = nhid // world
with torch.no_grad():
    par.conv1.W.weight.copy_(W1[rank * blk:(rank + 1) * blk])
    par.conv_out.W.weight.copy_(W2)
got = par(X)
loss = Fn.nll_loss(got, y)
loss.backward()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
errs = torch.tensor([rel(got, want), abs(loss.item() - wloss.item()),
                     rel(par.conv1.W.weight.grad, W1.grad[rank * blk:(rank + 1) * blk]),
                     rel(par.conv_out.W.weight.grad, W2.grad)], device=dev)
if world > 1:
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "out_err": errs[0].item(), "loss_err": errs[1].item(),
                      "grad_w1_block_err": errs[2].item(), "grad_w2_err": errs[3].item()}), flush=True)
if world > 1:
    dist.destroy_process_group()
