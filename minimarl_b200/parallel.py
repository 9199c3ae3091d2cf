"""The two multi-GPU pieces of the hot path: env sharding and the gradient exchange.

Env instances are independent, so they are partitioned by contiguous blocks, one process per GPU, with
no data-path collective.  The only exchange step is the learner's gradient all-reduce: one flattened
fp32 bucket (83 KB MAPPO / 260 KB QMIX), summed with NCCL over NVLink and divided by the world size --
the semantics of the reference's (dead) ``average_gradients`` (offpolicy/utils/util.py:148-153).
The collective is latency-bound at this size; it is issued once per optimiser step, before clipping.
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous block [lo, hi) of ``n_total`` env instances owned by ``rank`` (remainder to low ranks)."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_grads(params, group=None):
    """In-place average of ``p.grad`` over all ranks, one bucket, one collective.  Parameters without a
    gradient (e.g. the reference's unused ``fc_h`` duplicates, Q6) stay out of the bucket on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel() * flat.element_size()


def allreduce_stats(values, group=None):
    """Sum a small vector of scalars (advantage-normalisation sums, ValueNorm batch moments, PER min / sum /
    len) over ranks so that sharded runs reproduce the single-process formulas."""
    t = values if torch.is_tensor(values) else torch.tensor(values, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
