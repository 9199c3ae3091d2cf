"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (env sharding, gradient bucket all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minimarl_b200.parallel import (GradBucket, allreduce_mean_grads, allreduce_stats, per_global_weights,
                                    shard_range)


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
    # ---- flat bucket: p.grad are views, backward lands in the bucket, one collective, extras summed alongside --------
    torch.manual_seed(0)
    net2 = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    unused2 = torch.nn.Parameter(torch.ones(4))
    bucket = GradBucket(list(net2.parameters()) + [unused2], n_extra=2)
    bucket.zero()
    net2(x).sum().backward()
    ok = ok and bucket.attached()                               # autograd accumulated in place into the views
    bucket.extra.copy_(torch.tensor([float(rank + 1), 1.0]))
    nbytes2 = allreduce_mean_grads(bucket.params, bucket=bucket)
    got2 = torch.cat([p.grad.reshape(-1) for p in net2.parameters()])
    ok = ok and torch.allclose(got2, want, atol=1e-6) and bucket.extra.tolist() == [3.0, 2.0]
    ok = ok and float(unused2.grad.abs().sum()) == 0.0 and nbytes2 == bucket.flat.numel() * 4
    torch.optim.SGD(bucket.params, lr=0.1).zero_grad(set_to_none=True)          # views dropped -> falls back to packing
    ok = ok and not bucket.attached()
    # ---- sharded PER: weights under the global normaliser equal the single-process formula on the union ---------------
    g = torch.Generator().manual_seed(5)
    prios = torch.rand(10, dtype=torch.float64, generator=g) + 0.05             # 10 episodes, 6 on rank 0, 4 on rank 1
    mine = prios[:6] if rank == 0 else prios[6:]
    beta = 0.4
    w = per_global_weights(mine[:3], mine.min(), mine.sum(), len(mine), beta)
    p_min = prios.min() / prios.sum()
    want_w = (mine[:3] / prios.sum() * 10) ** (-beta) / (p_min * 10) ** (-beta)
    ok = ok and torch.allclose(w, want_w, rtol=1e-12)
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
