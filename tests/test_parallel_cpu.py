"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (env sharding, gradient bucket all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minimarl_b200.parallel import allreduce_mean_grads, allreduce_stats, shard_range


def test_shard_range_partitions_everything():
    for n, w in [(65536, 8), (10, 3), (7, 8), (16384, 1)]:
        blocks = [shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    unused = torch.nn.Parameter(torch.ones(4))                  # like the reference's unused fc_h: no grad
    params = list(net.parameters()) + [unused]
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    nbytes = allreduce_mean_grads(params)
    gathered = [torch.zeros_like(torch.cat([g.reshape(-1) for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
    want = sum(gathered) / world
    got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    stats = allreduce_stats([float(rank + 1), 1.0])
    ok = torch.allclose(got, want, atol=1e-6) and unused.grad is None and nbytes == got.numel() * 4
    ok = ok and stats.tolist() == [3.0, 2.0]
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_bucket_allreduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(0) and out.get(1)
