"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2: batch sharding, the flat gradient
buffer and its single all-reduce reproduce the single-process full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgat.parallel import FlatParams, shard_range


def test_shard_range_partitions_batch():
    for n in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _model()
    flat = FlatParams(model.named_parameters())
    flat.broadcast_params(0)
    torch.manual_seed(1)
    x, y = torch.rand(10, 6), torch.rand(10, 3)
    lo, hi = shard_range(10, rank, world)
    flat.zero_grad()
    # per-rank loss is a SUM over its samples so that the all-reduced sum equals the full-batch sum
    ((model(x[lo:hi]) - y[lo:hi]) ** 2).sum().backward()
    # gradients were accumulated into the flat buffer through the views
    assert all(p.grad.data_ptr() >= flat.grad.data_ptr() for p in model.parameters())
    w = flat.all_reduce_grads()
    assert w == world
    if rank == 0:
        torch.save(flat.grad.clone(), out)
    dist.destroy_process_group()


def test_flat_allreduce_matches_full_batch(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "grad.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    model = _model()
    flat = FlatParams(model.named_parameters())
    torch.manual_seed(1)
    x, y = torch.rand(10, 6), torch.rand(10, 3)
    ((model(x) - y) ** 2).sum().backward()
    torch.testing.assert_close(got, flat.grad, rtol=1e-6, atol=1e-6)


def test_flat_params_keep_state_dict_and_views():
    model = _model()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    flat = FlatParams(model.named_parameters())
    for k, v in model.state_dict().items():
        torch.testing.assert_close(v, before[k])
    flat.param.add_(1.0)  # an optimizer step on the flat buffer is visible through every parameter
    for k, v in model.state_dict().items():
        torch.testing.assert_close(v, before[k] + 1.0)
