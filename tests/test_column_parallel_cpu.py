"""The collective of the feature-column-parallel model (convs._GatherColumns: all_gather of column blocks whose
consumers are replicated) on CPU with gloo, world size 2 and 3: forward = the concatenated matrix, backward = the
local column slice of the incoming gradient, no reduction.  The aggregation around it is CUDA-only and is covered by
tests/test_convs.py::test_column_parallel_hgnn_matches_single_gpu."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hypergef_b200.convs import _GatherColumns


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)                     # the same full matrices on every rank
        N, cols = 11, 3
        full = torch.randn(N, cols * world, generator=g, dtype=torch.float64)
        Wout = torch.randn(cols * world, 5, generator=g, dtype=torch.float64)
        local = full[:, rank * cols:(rank + 1) * cols].clone().requires_grad_(True)
        gathered = _GatherColumns.apply(local, None)
        loss = (torch.tanh(gathered) @ Wout).pow(2).sum()        # a replicated consumer: the same loss on every rank
        loss.backward()
        ref = full.clone().requires_grad_(True)
        (torch.tanh(ref) @ Wout).pow(2).sum().backward()
        torch.save(dict(ok_fwd=torch.equal(gathered.detach(), full),
                        grad_err=float((local.grad - ref.grad[:, rank * cols:(rank + 1) * cols]).abs().max())),
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_columns_forward_and_gradient(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(tmp_path / f"r{r}.pt")
        assert res["ok_fwd"] and res["grad_err"] < 1e-12, (r, res)
